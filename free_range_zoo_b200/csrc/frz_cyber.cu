// Fused cybersecurity environment step for sm_100a.
//
// One launch = for every environment: attacker / defender action decode -> movement -> presence -> subnetwork state
// transition -> rewards -> num_moves / truncation -> reward accumulation -> observations -> task counts.  Replaces,
// with identical results, the reference's envs/cybersecurity/env/cybersecurity.py:296-526,
// env/transitions/{movement,presence,subnetwork}.py and utils/env.py:215-237.
//
// Mapping: one thread per environment (state is 24 B .. a few hundred bytes).  Rows of the [B, *] arrays are
// contiguous per environment, so the 32 rows a warp touches form one contiguous span and every sector fetched is
// fully used (L1 merges the strided per-thread accesses of a warp).
//
// The only transcendental on the path is tanh((patches - attacks) / T) (subnetwork.py:54) whose result is compared
// with a uniform draw.  Because few agents act on a node, the score takes at most 2^(Att+D) distinct values: they are
// tabulated on the host with the reference's own torch.tanh and indexed here by the set of agents acting on the node,
// so the threshold compare is bit-exact.
#include "frz_common.cuh"

namespace frz {
namespace {

constexpr int kCyThreads = 128;
enum CyMode { kCyStep = 0, kCyRefresh = 1 };

__global__ void __launch_bounds__(kCyThreads)
cyber_step_kernel(const __grid_constant__ FrzCyberParams p, const __grid_constant__ FrzCyberBuffers io, const int B,
                  const int mode) {
  const int N = p.num_nodes, n_att = p.num_attackers, n_def = p.num_defenders, n_agents = n_att + n_def;
  FrzControl* control = io.control;
  const uint64_t step = control->step;
  const uint32_t alive_prev = control->alive;
  const Philox philox(control->seed);
  const bool skip = (mode == kCyStep) && ((alive_prev & 3u) != 3u);  // utils/env.py:212
  const bool show_bad = p.flags & FRZ_CY_SHOW_BAD_ACTIONS;
  const bool stochastic = p.flags & FRZ_CY_STOCHASTIC_STATE;
  unsigned alive_bits = 0, faults = 0;

  if (!skip) {
    for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < B; env += gridDim.x * blockDim.x) {
      int32_t* state = io.network_state + size_t(env) * N;
      int32_t* location = io.location + size_t(env) * n_def;
      uint8_t* presence = io.presence + size_t(env) * n_agents;
      const size_t agent_row = size_t(env) * n_agents;

      uint32_t present = 0;  // bit a = agent a is present (attackers first)
      for (int a = 0; a < n_agents; ++a) present |= uint32_t(presence[a] != 0) << a;

      if (mode == kCyStep) {
        const int2* actions = reinterpret_cast<const int2*>(io.actions) + agent_row;
        const uint32_t env_lo = uint32_t(p.env_offset + env), env_hi = uint32_t(uint64_t(p.env_offset + env) >> 32);

        // ---------------------------------------------------------------- decode (cybersecurity.py:317-384)
        // acting[a] = node the agent attacks / patches this step, or -1
        uint32_t patching = 0, monitoring = 0, moving = 0;
        int8_t acting[FRZ_MAX_AGENTS];
        int32_t move_to[FRZ_MAX_AGENTS];
#pragma unroll 1
        for (int a = 0; a < n_agents; ++a) {
          const int2 act = actions[a];
          acting[a] = -1;
          move_to[a] = 0;
          if (!show_bad && !((present >> a) & 1u) && act.y != -1) faults |= FRZ_FAULT_ABSENT_ACTED;  // :345,362
          if (a < n_att) {
            if (act.y == 0) {  // attack node act.x
              if (act.x < 0 || act.x >= N) faults |= FRZ_FAULT_BAD_NODE;  // :341
              else acting[a] = int8_t(act.x);
            }
          } else {
            const int d = a - n_att;
            const int loc = location[d];
            if (act.y == 0) {  // move to node act.x
              if (act.x < 0 || act.x >= N) faults |= FRZ_FAULT_BAD_NODE;  // :357
              moving |= 1u << a;
              move_to[a] = act.x;
            } else if (act.y == -2 && loc != -1) {  // patch the node the defender stands on (pre-move) :354
              patching |= 1u << a;
              if (loc >= 0 && loc < N) acting[a] = int8_t(loc);
            }
            if (act.y == -3) monitoring |= 1u << a;
          }
        }

        // ---------------------------------------------------------------- movement.py:17-32 + presence.py:35-60
        uint32_t now_present = 0;
#pragma unroll 1
        for (int a = 0; a < n_agents; ++a) {
          float r;
          if (io.agent_uniforms != nullptr) {
            r = io.agent_uniforms[agent_row + a];
          } else {
            const uint4 bits = philox(env_lo, uint32_t(step), 0x80000000u | uint32_t(a >> 2), uint32_t(step >> 32) ^ env_hi);
            const uint32_t word = (a & 3) == 0 ? bits.x : (a & 3) == 1 ? bits.y : (a & 3) == 2 ? bits.z : bits.w;
            r = u01(word);
          }
          const bool was = (present >> a) & 1u;
          const bool returning = !was && r < p.returns[a];
          const bool leaving = was && r >= p.persist[a];
          const bool is = (was || returning) && !leaving;
          now_present |= uint32_t(is) << a;
          presence[a] = is;
          if (a >= n_att) {
            const int d = a - n_att;
            int loc = location[d];
            if ((moving >> a) & 1u) loc = move_to[a];
            if (returning) loc = -1;  // returning defenders start at the home node
            location[d] = loc;
          }
        }

        // ---------------------------------------------------------------- subnetwork.py:40-72 + rewards :395-409
        float network_reward = 0.f;
#pragma unroll 1
        for (int node = 0; node < N; ++node) {
          uint32_t actors = 0;
          float attacks = 0.f, patches = 0.f;
          for (int a = 0; a < n_agents; ++a) {
            if (acting[a] == node) {
              actors |= 1u << a;
              if (a < n_att) attacks = __fadd_rn(attacks, p.power[a]); else patches = __fadd_rn(patches, p.power[a]);
            }
          }
          float score;
          if (p.lut_bits > 0) score = io.score_lut[actors];
          else score = tanhf(__fdiv_rn(__fadd_rn(patches, -attacks), p.temperature));
          bool better = score > 0.f, worse = score < 0.f;
          if (stochastic) {
            float r;
            if (io.network_uniforms != nullptr) {
              r = io.network_uniforms[size_t(env) * N + node];
            } else {
              const uint4 bits = philox(env_lo, uint32_t(step), uint32_t(node >> 2), uint32_t(step >> 32) ^ env_hi);
              const uint32_t word = (node & 3) == 0 ? bits.x : (node & 3) == 1 ? bits.y : (node & 3) == 2 ? bits.z : bits.w;
              r = u01(word);
            }
            const bool within = fabsf(score) <= r;  // larger |score| => LESS likely, as in the reference
            better = better && within;
            worse = worse && within;
          }
          int s = state[node] - int(better) + int(worse);
          s = min(max(s, 0), p.num_states - 1);
          state[node] = s;
          network_reward = __fadd_rn(network_reward, __fmul_rn(p.state_rewards[s], p.criticality[node]));
          reinterpret_cast<int2*>(io.task_obs)[size_t(env) * N + node] = make_int2(s, int(p.criticality[node]));
        }

        const int moves = io.num_moves[env] + 1;
        const bool truncated = moves >= p.max_steps;
        io.num_moves[env] = moves;
        io.truncated[env] = truncated;
        alive_bits |= 1u | (truncated ? 0u : 2u);  // cybersecurity never terminates (:299)
#pragma unroll 1
        for (int a = 0; a < n_agents; ++a) {
          float reward = ((patching >> a) & 1u) ? p.patch_reward : 0.f;  // :376 (the bad-action branch :379-381 is dead)
          reward = __fadd_rn(reward, a < n_att ? -network_reward : network_reward);
          io.rewards[agent_row + a] = reward;
          io.cumulative_rewards[agent_row + a] = __fadd_rn(io.cumulative_rewards[agent_row + a], reward);
          if (a >= n_att) io.monitored[size_t(env) * n_def + (a - n_att)] = (monitoring >> a) & 1u;
        }
        present = now_present;
      } else {
        for (int node = 0; node < N; ++node)
          reinterpret_cast<int2*>(io.task_obs)[size_t(env) * N + node] = make_int2(state[node], int(p.criticality[node]));
      }

      // ------------------------------------------------------------------ update_actions / update_observations
      io.env_task_count[env] = N;
      for (int a = 0; a < n_agents; ++a) {
        const bool is = (present >> a) & 1u;
        io.agent_task_count[agent_row + a] = is ? N : 0;
        if (a < n_att) {
          reinterpret_cast<float2*>(io.attacker_self)[size_t(env) * n_att + a] = make_float2(p.power[a], is ? 1.f : 0.f);
        } else {
          const int d = a - n_att;
          float* out = io.defender_self + (size_t(env) * n_def + d) * 3;
          out[0] = p.power[a];
          out[1] = is ? 1.f : 0.f;
          out[2] = float(location[d]);
        }
      }
    }
  }
  finish_launch(control, alive_bits, faults, 0u,
                skip ? kPublishNothing : (mode == kCyStep ? kPublishStep : kPublishRefresh));
}

__global__ void cyber_restore_kernel(const FrzCyberParams p, const FrzCyberBuffers io, const int B,
                                     const uint8_t* __restrict__ env_mask) {
  const int N = p.num_nodes, n_def = p.num_defenders, n_agents = p.num_attackers + p.num_defenders;
  for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < B; env += gridDim.x * blockDim.x) {
    if (env_mask != nullptr && !env_mask[env]) continue;
    for (int i = 0; i < N; ++i) io.network_state[size_t(env) * N + i] = io.init_network_state[size_t(env) * N + i];
    for (int i = 0; i < n_def; ++i) {
      io.location[size_t(env) * n_def + i] = io.init_location[size_t(env) * n_def + i];
      io.monitored[size_t(env) * n_def + i] = 0;  // actions are re-initialised to -2 (cybersecurity.py:233-236)
    }
    for (int i = 0; i < n_agents; ++i) {
      io.presence[size_t(env) * n_agents + i] = io.init_presence[size_t(env) * n_agents + i];
      io.rewards[size_t(env) * n_agents + i] = 0.f;
      io.cumulative_rewards[size_t(env) * n_agents + i] = 0.f;
    }
    io.terminated[env] = 0;
    io.truncated[env] = 0;
    io.num_moves[env] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) io.control->alive = 3u;
}

// Uniform over the legal choices of spaces/actions.py:11-99:
//   attacker: [attack node 0..n-1, noop]           defender: [move 0..n-1, noop, (patch), monitor]
// with n = N if the agent is present (or show_bad_actions) else 0 (noop only); patch is offered unless
// show_bad_actions is off and the defender is at the home node.
__global__ void cyber_sample_kernel(const FrzCyberParams p, const FrzCyberBuffers io, const int B,
                                    const uint64_t sampler_seed) {
  const int n_att = p.num_attackers, n_agents = p.num_attackers + p.num_defenders;
  const Philox philox(sampler_seed);
  const uint64_t step = io.control->step;
  const bool show_bad = p.flags & FRZ_CY_SHOW_BAD_ACTIONS;
  const int total = B * n_agents;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int env = i / n_agents, a = i - env * n_agents;
    const int n = show_bad ? io.env_task_count[env] : io.agent_task_count[i];
    const uint64_t genv = uint64_t(p.env_offset + env);
    const uint4 r = philox(uint32_t(genv), uint32_t(step), 0xC0000000u | uint32_t(a), uint32_t(step >> 32) ^ uint32_t(genv >> 32));
    int choices = n + 1;
    bool can_patch = false;
    if (a >= n_att && n > 0) {
      can_patch = show_bad || io.location[size_t(env) * p.num_defenders + (a - n_att)] != -1;
      choices += can_patch ? 2 : 1;
    }
    const int k = min(int(u01(r.x) * float(choices)), choices - 1);
    int ident;
    if (k < n) ident = 0;
    else if (k == n) ident = -1;
    else if (k == n + 1 && can_patch) ident = -2;
    else ident = -3;
    reinterpret_cast<int2*>(const_cast<int32_t*>(io.actions))[i] = make_int2(k, ident);
  }
}

int cyber_validate(const FrzCyberParams* p, const FrzCyberBuffers* io, int B, const char* what) {
  if (p == nullptr || io == nullptr || io->control == nullptr || io->network_state == nullptr) {
    set_error("%s: NULL params / buffers", what);
    return FRZ_ERR_NULL;
  }
  const int n_agents = p->num_attackers + p->num_defenders;
  if (B <= 0 || p->num_nodes < 1 || p->num_nodes > FRZ_MAX_NODES || n_agents < 1 || n_agents > FRZ_MAX_AGENTS ||
      p->num_states < 1 || p->num_states > FRZ_MAX_NET_STATES ||
      (p->lut_bits != 0 && (p->lut_bits != n_agents || n_agents > FRZ_CY_MAX_LUT_BITS || io->score_lut == nullptr))) {
    set_error("%s: unsupported shape B=%d N=%d attackers=%d defenders=%d states=%d lut_bits=%d", what, B, p->num_nodes,
              p->num_attackers, p->num_defenders, p->num_states, p->lut_bits);
    return FRZ_ERR_SHAPE;
  }
  return FRZ_OK;
}

int cyber_launch(const FrzCyberParams* p, const FrzCyberBuffers* io, int B, int mode, void* stream) {
  const int work = (B + kCyThreads - 1) / kCyThreads;
  const int grid = persistent_grid(work, 8);
  cyber_step_kernel<<<grid, kCyThreads, 0, static_cast<cudaStream_t>(stream)>>>(*p, *io, B, mode);
  return check_launch("cyber_step_kernel");
}

}  // namespace
}  // namespace frz

extern "C" {

int frz_cyber_step(const FrzCyberParams* params, const FrzCyberBuffers* io, int32_t parallel_envs, void* stream) {
  const int status = frz::cyber_validate(params, io, parallel_envs, "frz_cyber_step");
  if (status != FRZ_OK) return status;
  if (io->actions == nullptr) {
    frz::set_error("frz_cyber_step: actions is NULL");
    return FRZ_ERR_NULL;
  }
  return frz::cyber_launch(params, io, parallel_envs, frz::kCyStep, stream);
}

int frz_cyber_refresh(const FrzCyberParams* params, const FrzCyberBuffers* io, int32_t parallel_envs, void* stream) {
  const int status = frz::cyber_validate(params, io, parallel_envs, "frz_cyber_refresh");
  if (status != FRZ_OK) return status;
  return frz::cyber_launch(params, io, parallel_envs, frz::kCyRefresh, stream);
}

int frz_cyber_reset(const FrzCyberParams* params, const FrzCyberBuffers* io, int32_t parallel_envs,
                    const uint8_t* env_mask, void* stream) {
  int status = frz::cyber_validate(params, io, parallel_envs, "frz_cyber_reset");
  if (status != FRZ_OK) return status;
  if (io->init_network_state == nullptr) {
    frz::set_error("frz_cyber_reset: initial state is NULL");
    return FRZ_ERR_NULL;
  }
  const int grid = frz::persistent_grid((parallel_envs + 255) / 256, 8);
  frz::cyber_restore_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*params, *io, parallel_envs, env_mask);
  status = frz::check_launch("cyber_restore_kernel");
  if (status != FRZ_OK) return status;
  return frz::cyber_launch(params, io, parallel_envs, frz::kCyRefresh, stream);
}

int frz_cyber_sample_actions(const FrzCyberParams* params, const FrzCyberBuffers* io, int32_t parallel_envs,
                             uint64_t sampler_seed, void* stream) {
  const int status = frz::cyber_validate(params, io, parallel_envs, "frz_cyber_sample_actions");
  if (status != FRZ_OK) return status;
  const int total = parallel_envs * (params->num_attackers + params->num_defenders);
  const int grid = frz::persistent_grid((total + 255) / 256, 8);
  frz::cyber_sample_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*params, *io, parallel_envs, sampler_seed);
  return frz::check_launch("cyber_sample_kernel");
}

}  // extern "C"
