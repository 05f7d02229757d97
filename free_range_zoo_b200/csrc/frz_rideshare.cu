// Fused rideshare environment step for sm_100a.
//
// One launch = for every environment: action decode -> movement -> passenger state (accept conflicts, picks) ->
// passenger exit -> passenger entry -> rewards -> num_moves / truncation -> reward accumulation -> observations ->
// task lists.  Replaces, with identical results, the reference's envs/rideshare/env/rideshare.py:249-467,
// env/transitions/{movement,passenger_state,passenger_exit,passenger_entry}.py and utils/env.py:215-237.
//
// The reference keeps all passengers of all environments in one flat table [N_total, 11] sorted by environment and
// re-sorts / compacts it with `unique`, boolean-mask indexing and a stable `argsort` every step.  Here environment b
// owns rows passengers[b, 0:count[b]] (same relative order).
//
// Mapping: a group of G lanes (8, 16 or 32, chosen from the table capacity and the number of drivers) owns one
// environment, so a warp steps 4, 2 or 1 environments at once.  The environment's rows are staged in shared memory
// (cp.async) for the whole step and edited in place; lane s of the group looks after rows s, s+G, ... (PPL of them)
// and driver s.  One table buffer per group: double buffering (prefetching the group's next environment) halved the
// resident warps and measured 8 % slower than letting twice as many warps cover the staging latency.  Task lists, conflict
// detection, compaction and the per-class "last row" reward quirk are ballots (cut to the group's lanes), match_any
// and popcounts on 64-bit row masks.
#include <algorithm>
#include <climits>
#include <math_constants.h>

#include "frz_common.cuh"
#include "frz_host.cuh"

namespace frz {
namespace {

constexpr int kRsThreads = 128;
constexpr int kCols = FRZ_RS_PASSENGER_COLUMNS;
enum RsMode { kRsStep = 0, kRsRefresh = 1, kRsEntryRefresh = 2 };
enum Col { cBatch = 0, cY, cX, cDestY, cDestX, cFare, cState, cAssoc, cEntered, cAccepted, cPicked };
constexpr int kRemoved = -7;  // scratch value of the state column: the row left the table in this step

// stay, N, E, S, W, NW, NE, SE, SW (transitions/movement.py:27-45)
__constant__ int kDirY[9] = {0, -1, 0, 1, 0, -1, -1, 1, 1};
__constant__ int kDirX[9] = {0, 0, 1, 0, -1, -1, 1, 1, -1};

__device__ __forceinline__ int select_bit64(uint64_t mask, int k) {
  const uint32_t lo = uint32_t(mask);
  const int in_lo = __popc(lo);
  return k < in_lo ? select_bit(lo, k) : 32 + select_bit(uint32_t(mask >> 32), k - in_lo);
}

// Squared Euclidean length.  The reference compares fp32 L2 norms of small integer vectors (movement.py:57-116,
// passenger_state.py:48-74); sqrt is strictly increasing on these exactly representable integers (neighbouring values are
// > 1e-3 apart, far above an fp32 ulp), so argmin, ties, "== 0" and "< 1e-6" give the same answers on the squares.
__device__ __forceinline__ int squared(int dy, int dx) { return dy * dy + dx * dx; }

// Row mask of one environment from per-lane predicates: bit (s + G * i) = predicate i of the group's lane s.  Every
// lane of the warp must call it (the ballots are warp-wide); each lane receives its own group's mask.  The group's
// byte / halfword of up to four ballots is packed with byte permutes (`pack` = the group's selector, see the kernel).
template <int G, int PPL>
__device__ __forceinline__ uint64_t group_rows(const bool (&pred)[PPL], int group_base, uint32_t pack) {
  uint32_t ballot[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) ballot[i] = __ballot_sync(kFullMask, pred[i]);
  if constexpr (G == 32) {
    return uint64_t(ballot[0]) | (PPL > 1 ? uint64_t(ballot[PPL - 1]) << 32 : 0);
  } else if constexpr (PPL == 1) {
    return uint64_t((ballot[0] >> group_base) & ((1u << G) - 1u));
  } else if constexpr (G == 16) {  // halfwords: rows (0, 1) -> low word, rows (2, 3) -> high word
    const uint32_t lo = __byte_perm(ballot[0], ballot[1], pack);
    const uint32_t hi = PPL > 2 ? __byte_perm(ballot[2], ballot[PPL - 1], pack) : 0u;
    return uint64_t(lo) | (uint64_t(hi) << 32);
  } else {  // G == 8, bytes: byte i of the result is the group's byte of ballot i
    static_assert(G == 8 && (PPL == 2 || PPL == 4), "8-lane groups hold 1, 2 or 4 rows per lane");
    const uint32_t lo = __byte_perm(ballot[0], ballot[1], pack);  // bytes 0, 1 valid
    if constexpr (PPL == 2) return uint64_t(lo & 0xffffu);
    const uint32_t hi = __byte_perm(ballot[2], ballot[PPL - 1], pack);
    return uint64_t(__byte_perm(lo, hi, 0x5410));
  }
}

template <int G>
__device__ __forceinline__ int group_min(int v) {
#pragma unroll
  for (int offset = G / 2; offset >= 1; offset >>= 1) v = min(v, __shfl_xor_sync(kFullMask, v, offset, G));
  return v;
}

// MODE is a template parameter so that the refresh / reset variants carry none of the step's code.
template <int G, int PPL, int MODE>
__global__ void __launch_bounds__(kRsThreads)
rideshare_step_kernel(const __grid_constant__ FrzRideshareParams p, const __grid_constant__ FrzRideshareBuffers io,
                      const int B, const uint8_t* __restrict__ entry_mask, const int batch_base) {
  static_assert(G * PPL <= 64, "row masks are 64 bits wide");
  extern __shared__ int smem[];
  constexpr int kGroupsPerWarp = 32 / G;
  constexpr int kGroupsPerCta = (kRsThreads / 32) * kGroupsPerWarp;
  const int lane = threadIdx.x & 31;
  // broadcast from lane 0: makes the warp index, hence the environment loop and every ballot, provably warp-uniform
  const int warp = __shfl_sync(kFullMask, int(threadIdx.x >> 5), 0);
  const int sub = lane % G, group_base = lane - sub;
  const int K = p.capacity, A = p.num_agents, S = p.schedule_rows;
  // this group's passenger-table buffer as a 32-bit shared-window address: rows are read / written as [table + 4 * index]
  const uint32_t table_bytes = uint32_t(K * kCols) * 4u;
  const uint32_t table = shared_address(smem) + uint32_t(warp * kGroupsPerWarp + lane / G) * table_bytes;
  const uint64_t rows_below = (uint64_t(1) << sub) - 1u;  // rows before this lane's first row
  // byte-permute selector that pulls this group's slice out of two ballots (group_rows): 16-lane groups take halfword
  // g of each, 8-lane groups byte g of each
  const uint32_t pack = (G == 16) ? ((group_base & 16) ? 0x7632u : 0x5410u)
                                  : (uint32_t(group_base >> 3) | ((4u + uint32_t(group_base >> 3)) << 4));

  FrzControl* control = io.control;
  const uint32_t alive_prev = control->alive;
  const uint32_t agents_with_tasks = control->agents_with_tasks;
  const bool skip = (MODE == kRsStep) && ((alive_prev & 3u) != 3u);  // utils/env.py:212
  const bool is_agent = sub < A;
  const bool fast = p.flags & FRZ_RS_FAST_TRAVEL, diagonal = p.flags & FRZ_RS_DIAGONAL_TRAVEL;
  const int directions = diagonal ? 9 : 5;
  // task-mask stores: work item -> (agent, four consecutive table rows)
  const bool wide_rows = (K & 3) == 0;  // every environment's table starts on a 16-byte boundary
  const int quads = (K + 3) >> 2;
  const uint32_t inverse_quads = (65536u + uint32_t(quads) - 1u) / uint32_t(quads);  // item / quads for item < 512
  unsigned alive_bits = 0, faults = 0, agent_bits = 0;

  if (!skip) {
    const int stride = gridDim.x * kGroupsPerCta;
    for (int env0 = (blockIdx.x * (kRsThreads / 32) + warp) * kGroupsPerWarp; env0 < B; env0 += stride) {
      const int env = min(env0 + lane / G, B - 1);
      const bool valid = env0 + lane / G < B;  // groups past the end of the batch replay the last environment, storing nothing
      const uint32_t agent_at = uint32_t(env) * uint32_t(A) + uint32_t(sub);
      int* const global_rows = io.passengers + uint32_t(env) * uint32_t(K * kCols);
      // stage the environment's live rows (cp.async); the other warps of the SM cover the latency
      const int n_before = min(io.env_task_count[env], K);
      {
        // 16-byte pieces while they lie inside the live rows (table bases are 16-byte aligned when K is a multiple of
        // four), single words for the remaining one to three
        const int words = n_before * kCols, quads_of_words = wide_rows ? words >> 2 : 0;
        for (int i = sub; i < quads_of_words; i += G) cp_async_16(table + 16u * i, global_rows + 4 * i);
        for (int i = 4 * quads_of_words + sub; i < words; i += G) cp_async_4(table + 4u * i, global_rows + i);
      }
      cp_async_commit();

      int agent_y = 0, agent_x = 0;
      int2 act = make_int2(0, -100);
      if (is_agent) {
        const int2 at = reinterpret_cast<const int2*>(io.agents)[agent_at];
        agent_y = at.x;
        agent_x = at.y;
        if (MODE == kRsStep) act = reinterpret_cast<const int2*>(io.actions)[agent_at];
      }
      const int t_now = io.num_moves[env];
      int n_kept = n_before, fare_won = 0;
      float move_cost = 0.f;
      const bool noop = act.y == -1, accept = act.y == 0, pick = act.y == 1, drop = act.y == 2;
      cp_async_wait_all();
      __syncwarp();

      // state / association columns of this lane's rows (kept in registers, refreshed after every edit phase)
      int state[PPL], assoc[PPL];
      bool present[PPL];
      const auto load_rows = [&](int count) {
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          const int r = sub + G * i;
          present[i] = r < count;
          state[i] = present[i] ? int(lds(table + 4u * uint32_t(r * kCols + cState))) : -1;
          assoc[i] = present[i] ? int(lds(table + 4u * uint32_t(r * kCols + cAssoc))) : -1;
        }
      };
      bool pred[PPL];

      if (MODE == kRsStep) {
        load_rows(n_before);
        // ---------------------------------------------------------------- task lists (rideshare.py:374-386)
#pragma unroll
        for (int i = 0; i < PPL; ++i) pred[i] = present[i] && state[i] == 0;
        const uint64_t unaccepted = group_rows<G, PPL>(pred, group_base, pack);
        uint64_t mine = 0;
        for (int a = 0; a < A; ++a) {
#pragma unroll
          for (int i = 0; i < PPL; ++i) pred[i] = present[i] && assoc[i] == a;
          const uint64_t own = group_rows<G, PPL>(pred, group_base, pack);
          if (sub == a) mine = unaccepted | own;
        }

        // ---------------------------------------------------------------- decode (rideshare.py:255-300)
        int target = -1;
        // the reference only resolves targets of agents that have a task in SOME environment (rideshare.py:276)
        if (is_agent && !noop && ((agents_with_tasks >> sub) & 1u)) {
          if (act.x >= 0 && act.x < __popcll(mine)) target = select_bit64(mine, act.x);
          else if (valid && (accept || pick || drop)) faults |= FRZ_FAULT_BAD_TASK_INDEX;
        }
        const bool has_vector = target >= 0 && (accept || pick || drop);
        int goal_y = 0, goal_x = 0, fare_target = 0;
        if (has_vector) {
          const uint32_t t = table + 4u * uint32_t(target * kCols);
          goal_y = int(lds(t + 4u * (drop ? cDestY : cY)));
          goal_x = int(lds(t + 4u * (drop ? cDestX : cX)));
          fare_target = int(lds(t + 4u * cFare));
        }

        // ---------------------------------------------------------------- movement (transitions/movement.py:57-116)
        int move_y = 0, move_x = 0;
        int distance2 = INT_MAX;  // squared distance agent -> goal before moving (passenger_state.py:48-49)
        if (has_vector) {
          distance2 = squared(agent_y - goal_y, agent_x - goal_x);
          if (fast) {
            move_y = goal_y - agent_y;
            move_x = goal_x - agent_x;
          } else {
            int best = INT_MAX;
            for (int d = 0; d < directions; ++d) {  // first argmin: strict <
              const int candidate = squared(agent_y + kDirY[d] - goal_y, agent_x + kDirX[d] - goal_x);
              if (candidate < best) {
                best = candidate;
                move_y = kDirY[d];
                move_x = kDirX[d];
              }
            }
          }
          move_cost = diagonal ? __fsqrt_rn(float(squared(move_y, move_x))) : float(abs(move_y) + abs(move_x));
        }
        agent_y += move_y;
        agent_x += move_x;
        __syncwarp();  // every lane has read its goal from the table
        // riding passengers travel with their driver; association -1 wraps to the last agent like the tensor index
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
          int driver = assoc[i] < 0 ? assoc[i] + A : assoc[i];
          driver = group_base + min(max(driver, 0), G - 1);
          const int dy = __shfl_sync(kFullMask, move_y, driver), dx = __shfl_sync(kFullMask, move_x, driver);
          if (present[i] && state[i] == 2 && (dy | dx) != 0) {
            const uint32_t t = table + 4u * uint32_t((sub + G * i) * kCols);
            sts(t + 4u * cY, uint32_t(int(lds(t + 4u * cY)) + dy));
            sts(t + 4u * cX, uint32_t(int(lds(t + 4u * cX)) + dx));
          }
        }

        // ---------------------------------------------------------------- accept conflicts (passenger_state.py:50-74)
        int claim = (accept && target >= 0) ? target : -1;
        {
          const unsigned key = claim >= 0 ? (unsigned(group_base) << 8 | unsigned(claim)) : (0x80000000u | unsigned(lane));
          const bool duplicated = __popc(__match_any_sync(kFullMask, key)) > 1;
          if (__any_sync(kFullMask, duplicated)) {
            // among ALL duplicated claims of the environment only the closest claimant (first on ties) survives
            const int contest = duplicated ? distance2 : INT_MAX;
            const int closest = group_min<G>(contest);
            const uint32_t tied = (__ballot_sync(kFullMask, duplicated && contest == closest) >> group_base);
            const int keeper = __ffs(G == 32 ? tied : (tied & ((1u << G) - 1u))) - 1;
            if (duplicated && sub != keeper) claim = -1;
          }
        }
        const int picked = (pick && target >= 0 && distance2 == 0) ? target : -1;   // passenger_state.py:89-92 (< 1e-6)
        const int dropped = (drop && target >= 0 && distance2 == 0) ? target : -1;  // passenger_exit.py:40-45 (== 0)
        fare_won = dropped >= 0 ? fare_target : 0;

        // ---------------------------------------------------------------- apply: the drivers edit the rows in place
        if (claim >= 0) {  // accept: all accepts land before any pick, like the reference's transition
          const uint32_t t = table + 4u * uint32_t(claim * kCols);
          sts(t + 4u * cState, 1u);
          sts(t + 4u * cAccepted, uint32_t(t_now));
          sts(t + 4u * cAssoc, uint32_t(sub));
        }
        __syncwarp();
        if (picked >= 0) {
          const uint32_t t = table + 4u * uint32_t(picked * kCols);
          sts(t + 4u * cState, 2u);
          sts(t + 4u * cPicked, uint32_t(t_now));
        }
        __syncwarp();
        if (dropped >= 0) sts(table + 4u * uint32_t(dropped * kCols + cState), uint32_t(kRemoved));  // after the picks
        __syncwarp();

        // ---------------------------------------------------------------- exit: order-preserving compaction
        if (__any_sync(kFullMask, dropped >= 0)) {
#pragma unroll
          for (int i = 0; i < PPL; ++i) {
            const int r = sub + G * i;
            pred[i] = r < n_before && int(lds(table + 4u * uint32_t(r * kCols + cState))) != kRemoved;
          }
          const uint64_t kept = group_rows<G, PPL>(pred, group_base, pack);
          n_kept = __popcll(kept);
          // rows move down to their rank among the kept rows; slab by slab (rows s + G * i for all lanes s), read then
          // write: a row lands either in an earlier slab (done) or on a row of this slab that was just read
#pragma unroll
          for (int i = 0; i < PPL; ++i) {
            const int r = sub + G * i;
            const int to = __popcll(kept & ((rows_below << (G * i)) | ((uint64_t(1) << (G * i)) - 1u)));
            const bool moves = pred[i] && to != r;
            int row[kCols];
            if (moves) {
#pragma unroll
              for (int c = 0; c < kCols; ++c) row[c] = int(lds(table + 4u * uint32_t(r * kCols + c)));
            }
            __syncwarp();
            if (moves) {
#pragma unroll
              for (int c = 0; c < kCols; ++c) sts(table + 4u * uint32_t(to * kCols + c), uint32_t(row[c]));
            }
            __syncwarp();
          }
        }
      }

      // ------------------------------------------------------------------ entry (passenger_entry.py:25-72)
      int n_rows = n_kept;
      const bool admits = MODE == kRsStep || (MODE == kRsEntryRefresh && (entry_mask == nullptr || entry_mask[env]));
      {
        const int t_entry = (MODE == kRsStep) ? t_now + 1 : t_now;  // rideshare.py:307 vs :212
        // rows of the time-sorted schedule that enter at t_entry: [schedule_index[t], schedule_index[t + 1])
        int lo = S, hi = S;
        if (admits && t_entry >= 0 && t_entry <= p.schedule_horizon) {
          lo = io.schedule_index[t_entry];
          hi = io.schedule_index[t_entry + 1];
        }
        if (__any_sync(kFullMask, lo < hi)) {  // (environments of one warp may be at different steps after a partial reset)
          const int64_t global_env = p.env_offset + env;
          constexpr uint32_t lanes = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
          for (int base = 0; __any_sync(kFullMask, lo + base < hi); base += G) {
            const int r = lo + base + sub;
            const bool now = r < hi;
            const int* s = io.schedule + r * 7;
            const int batch = now ? s[1] : 0;
            const bool enters = now && (batch == -1 || batch == global_env);
            const uint32_t entering = (__ballot_sync(kFullMask, enters) >> group_base) & lanes;
            const int slot = n_rows + __popc(entering & ((1u << sub) - 1u));
            if (enters) {
              if (slot < K) {
                const uint32_t out = table + 4u * uint32_t(slot * kCols);
                sts(out + 4u * cBatch, uint32_t(batch_base + env));  // index in the caller's batch (a slice starts at batch_base)
                sts(out + 4u * cY, uint32_t(s[2]));
                sts(out + 4u * cX, uint32_t(s[3]));
                sts(out + 4u * cDestY, uint32_t(s[4]));
                sts(out + 4u * cDestX, uint32_t(s[5]));
                sts(out + 4u * cFare, uint32_t(s[6]));
                sts(out + 4u * cState, 0u);
                sts(out + 4u * cAssoc, uint32_t(-1));
                sts(out + 4u * cEntered, uint32_t(t_entry));
                sts(out + 4u * cAccepted, uint32_t(-1));
                sts(out + 4u * cPicked, uint32_t(-1));
              } else if (valid) {
                faults |= FRZ_FAULT_TABLE_FULL;
              }
            }
            n_rows = min(n_rows + __popc(entering), K);
          }
        }
      }
      __syncwarp();

      // ------------------------------------------------------------------ the new table
      load_rows(n_rows);
#pragma unroll
      for (int i = 0; i < PPL; ++i) {
        const int r = sub + G * i;
        // task observation row (rideshare.py:398-416) + padding of rows that just became free
        if (valid && r < K && (present[i] || r < n_before || MODE != kRsStep)) {
          int4 head = make_int4(FRZ_PAD, FRZ_PAD, FRZ_PAD, FRZ_PAD), tail = head;
          if (present[i]) {
            const uint32_t t = table + 4u * uint32_t(r * kCols);
            head = make_int4(int(lds(t + 4u * cY)), int(lds(t + 4u * cX)), int(lds(t + 4u * cDestY)), int(lds(t + 4u * cDestX)));
            tail = make_int4(state[i] == 1 ? assoc[i] : FRZ_PAD, state[i] == 2 ? assoc[i] : FRZ_PAD, int(lds(t + 4u * cFare)),
                             int(lds(t + 4u * cEntered)));
          }
          int4* out = reinterpret_cast<int4*>(io.task_obs) + (uint32_t(env) * uint32_t(K) + uint32_t(r)) * 2u;
          out[0] = head;
          out[1] = tail;
        }
      }

      // row masks by passenger state, per-agent counts; lane a of the group keeps agent a's
      uint64_t in_state[3];
#pragma unroll
      for (int cls = 0; cls < 3; ++cls) {
#pragma unroll
        for (int i = 0; i < PPL; ++i) pred[i] = present[i] && state[i] == cls;
        in_state[cls] = group_rows<G, PPL>(pred, group_base, pack);
      }
      int associated = 0, n_accepted = 0, n_riding = 0, n_tasks = 0;
      uint64_t members_mine = 0;
      for (int a = 0; a < A; ++a) {
#pragma unroll
        for (int i = 0; i < PPL; ++i) pred[i] = present[i] && assoc[i] == a;
        const uint64_t own = group_rows<G, PPL>(pred, group_base, pack);
        if (sub == a) {
          associated = __popcll(own);
          n_accepted = __popcll(own & in_state[1]);
          n_riding = __popcll(own & in_state[2]);
          members_mine = in_state[0] | own;
          n_tasks = __popcll(members_mine);
        }
      }
      // task mask [A, K] bytes: work item (agent, four consecutive rows) -> one 4-byte store when K is a multiple of 4
      {
        const bool aligned = (K & 3) == 0;
        const int items = aligned ? A * quads : 0;
        uint32_t* const mask_words = reinterpret_cast<uint32_t*>(io.task_mask);
        const uint32_t mask_at = uint32_t(env) * uint32_t(A * quads);
        for (int first = 0; first < items; first += G) {
          const int item = first + sub;
          const int a = int((uint32_t(item) * inverse_quads) >> 16), q = item - a * quads;
          const int source = group_base + min(a, G - 1);
          const uint32_t lo = __shfl_sync(kFullMask, uint32_t(members_mine), source);
          const uint32_t hi = __shfl_sync(kFullMask, uint32_t(members_mine >> 32), source);
          const uint32_t nibble = ((q < 8 ? lo : hi) >> (4 * (q & 7))) & 0xfu;
          if (valid && item < items) mask_words[mask_at + item] = (nibble * 0x00204081u) & 0x01010101u;
        }
        if (!aligned) {
          for (int a = 0; a < A; ++a) {
            const uint64_t members = (uint64_t(__shfl_sync(kFullMask, uint32_t(members_mine >> 32), group_base + a)) << 32) |
                                     __shfl_sync(kFullMask, uint32_t(members_mine), group_base + a);
            uint8_t* mask_row = io.task_mask + (uint32_t(env) * uint32_t(A) + uint32_t(a)) * uint32_t(K);
#pragma unroll
            for (int i = 0; i < PPL; ++i) {
              const int r = sub + G * i;
              if (valid && r < K) mask_row[r] = (members >> r) & 1u;
            }
          }
        }
      }

      if (MODE == kRsStep) {
        // ---------------------------------------------------------------- rewards (rideshare.py:309-363)
        float shared = 0.f;
        if (p.flags & FRZ_RS_WAITING_COSTS) {
          // `global[idx] += v` with duplicate indices keeps ONE write per statement: the last row of the class
          int elapsed_unaccepted = 0;
#pragma unroll
          for (int cls = 0; cls < 3; ++cls) {
            if (in_state[cls]) {
              const int last = 63 - __clzll(in_state[cls]);
              const int when = int(lds(table + 4u * uint32_t(last * kCols + (cls == 0 ? cEntered : (cls == 1 ? cAccepted : cPicked)))));
              const int elapsed = t_now - when;
              shared = __fadd_rn(shared, __fmul_rn(elapsed >= p.wait_limit[cls] ? 1.f : 0.f, p.general_wait_cost));
              if (cls == 0) elapsed_unaccepted = elapsed;
            }
          }
          if (in_state[0])
            shared = __fadd_rn(shared, __fmul_rn(elapsed_unaccepted >= p.long_wait_time ? 1.f : 0.f, p.long_wait_cost));
          const int free_slots = A * p.pool_limit - n_rows;
          const float unserved = __fmul_rn(__popcll(in_state[0]) >= free_slots ? 1.f : 0.f, -0.5f);
          shared = __fadd_rn(shared, __fmul_rn(unserved, float(free_slots)));
        }
        const int moves = t_now + 1;
        const bool truncated = moves >= p.max_steps;
        if (is_agent && valid) {
          float reward = associated > p.pool_limit ? p.pool_limit_cost : 0.f;
          reward = __fadd_rn(reward, __fmul_rn(noop ? 1.f : 0.f, p.noop_cost));
          reward = __fadd_rn(reward, __fmul_rn(accept ? 1.f : 0.f, p.accept_cost));
          reward = __fadd_rn(reward, fare_won > 0 ? __fadd_rn(float(fare_won), -p.drop_cost) : 0.f);
          float move_reward = __fmul_rn(move_cost, p.move_cost);
          if (p.flags & FRZ_RS_VARIABLE_MOVE_COST) move_reward = __fdiv_rn(move_reward, float(associated + 1));
          reward = __fadd_rn(reward, move_reward);
          reward = __fadd_rn(reward, shared);
          io.rewards[agent_at] = reward;
          io.cumulative_rewards[agent_at] = __fadd_rn(io.cumulative_rewards[agent_at], reward);
          reinterpret_cast<int2*>(io.agents)[agent_at] = make_int2(agent_y, agent_x);
        }
        if (sub == 0 && valid) {
          io.num_moves[env] = moves;
          io.truncated[env] = truncated;
        }
        if (valid) alive_bits |= 1u | (truncated ? 0u : 2u);  // rideshare never terminates (rideshare.py:252)
      }

      // ------------------------------------------------------------------ publish
      if (valid) {
        if (is_agent) {
          io.agent_task_count[agent_at] = n_tasks;
          if (n_tasks > 0) agent_bits |= 1u << sub;
          reinterpret_cast<int4*>(io.self_obs)[agent_at] = make_int4(agent_y, agent_x, n_accepted, n_riding);
        }
        if (sub == 0) io.env_task_count[env] = n_rows;
        if (admits) {
          const int words = n_rows * kCols, quads_of_words = wide_rows ? words >> 2 : 0;
          for (int i = sub; i < quads_of_words; i += G) {
            const uint4 piece = lds_v4(table + 16u * i);
            reinterpret_cast<int4*>(global_rows)[i] = make_int4(int(piece.x), int(piece.y), int(piece.z), int(piece.w));
          }
          for (int i = 4 * quads_of_words + sub; i < words; i += G) global_rows[i] = int(lds(table + 4u * i));
        }
      }
      __syncwarp();
    }
  }
  finish_launch(control, alive_bits, faults, agent_bits,
                skip ? kPublishNothing : (MODE == kRsStep ? kPublishStep : kPublishRefresh));
}

__global__ void rideshare_restore_kernel(const FrzRideshareParams p, const FrzRideshareBuffers io, const int B,
                                         const uint8_t* __restrict__ env_mask) {
  const int A = p.num_agents, K = p.capacity;
  // items per environment: table words, or driver coordinates when the table is shorter than those (a one-row
  // schedule with many drivers) -- every driver and every reward must be reached
  const int table_words = K * kCols;
  const int per_env = table_words > 2 * A ? table_words : 2 * A;
  const size_t total = size_t(B) * per_env;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
    const int env = int(i / per_env), j = int(i % per_env);
    if (env_mask != nullptr && !env_mask[env]) continue;
    if (j < table_words) io.passengers[size_t(env) * table_words + j] = io.init_passengers[size_t(env) * table_words + j];
    if (j < 2 * A) io.agents[size_t(env) * 2 * A + j] = io.init_agents[size_t(env) * 2 * A + j];
    if (j < A) {
      io.rewards[size_t(env) * A + j] = 0.f;
      io.cumulative_rewards[size_t(env) * A + j] = 0.f;
    }
    if (j == 0) {
      io.env_task_count[env] = io.init_count[env];
      io.terminated[env] = 0;
      io.truncated[env] = 0;
      io.num_moves[env] = 0;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) io.control->alive = 3u;
}

// Uniform over [task 0 .. task n-1, noop]; the action id of a task is the passenger's state (0 accept / 1 pick /
// 2 drop), which is the only id the reference's action space offers for it (spaces/actions.py:10-50).
__global__ void rideshare_sample_kernel(const FrzRideshareParams p, const FrzRideshareBuffers io, const int B,
                                        const uint64_t sampler_seed) {
  const int A = p.num_agents, K = p.capacity;
  const Philox philox(sampler_seed);
  const uint64_t step = io.control->step;
  const int total = B * A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int env = i / A, agent = i - env * A;
    const int n = io.agent_task_count[i];
    const uint64_t genv = uint64_t(p.env_offset + env);
    const uint4 r = philox(uint32_t(genv), uint32_t(step), 0xC0000000u | uint32_t(agent), uint32_t(step >> 32) ^ uint32_t(genv >> 32));
    const int k = min(int(u01(r.x) * float(n + 1)), n);
    int ident = -1;
    if (k < n) {
      const uint8_t* mask = io.task_mask + size_t(i) * K;
      int seen = 0;
      for (int row = 0; row < K; ++row) {
        if (mask[row]) {
          if (seen == k) {
            const int* obs = io.task_obs + (size_t(env) * K + row) * FRZ_RS_TASK_COLUMNS;
            ident = obs[5] != FRZ_PAD ? 2 : (obs[4] != FRZ_PAD ? 1 : 0);
            break;
          }
          ++seen;
        }
      }
    }
    reinterpret_cast<int2*>(const_cast<int32_t*>(io.actions))[i] = make_int2(k, ident);
  }
}

int rideshare_validate(const FrzRideshareParams* p, const FrzRideshareBuffers* io, int B, const char* what) {
  if (p == nullptr || io == nullptr || io->control == nullptr || io->passengers == nullptr || io->agents == nullptr ||
      (p->schedule_rows > 0 && (io->schedule == nullptr || io->schedule_index == nullptr))) {
    set_error("%s: NULL params / buffers", what);
    return FRZ_ERR_NULL;
  }
  if (B <= 0 || p->num_agents < 1 || p->num_agents > FRZ_MAX_AGENTS || p->capacity < 1 ||
      p->capacity > FRZ_MAX_PASSENGERS || p->schedule_rows < 0 ||
      uint64_t(B) * uint64_t(p->capacity) * FRZ_RS_PASSENGER_COLUMNS >= (1ull << 32)) {  // 32-bit element indices
    set_error("%s: unsupported shape B=%d agents=%d capacity=%d schedule=%d", what, B, p->num_agents, p->capacity,
              p->schedule_rows);
    return FRZ_ERR_SHAPE;
  }
  return FRZ_OK;
}

// Lanes per environment and rows per lane: the smallest group that holds the drivers and keeps at most 4 rows per lane.
template <int G, int PPL, int MODE>
int rideshare_launch_geometry(const FrzRideshareParams* p, const FrzRideshareBuffers* io, int B, cudaStream_t s,
                              const uint8_t* entry_mask, int batch_base) {
  constexpr int groups_per_cta = (kRsThreads / 32) * (32 / G);
  const size_t smem = size_t(groups_per_cta) * p->capacity * kCols * sizeof(int);  // one table buffer per group
  auto kernel = rideshare_step_kernel<G, PPL, MODE>;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)) != cudaSuccess)
    return check_launch("rideshare shared memory");
  const int ctas_per_sm = resident_ctas(kernel, kRsThreads, smem);
  const int grid = persistent_grid((B + groups_per_cta - 1) / groups_per_cta, ctas_per_sm);
  kernel<<<grid, kRsThreads, smem, s>>>(*p, *io, B, entry_mask, batch_base);
  return check_launch("rideshare_step_kernel");
}

template <int MODE>
int rideshare_launch_mode(const FrzRideshareParams* p, const FrzRideshareBuffers* io, int B, cudaStream_t s,
                          const uint8_t* entry_mask, int batch_base) {
  const int K = p->capacity, A = p->num_agents;
  if (A <= 8 && K <= 8) return rideshare_launch_geometry<8, 1, MODE>(p, io, B, s, entry_mask, batch_base);
  if (A <= 8 && K <= 16) return rideshare_launch_geometry<8, 2, MODE>(p, io, B, s, entry_mask, batch_base);
  if (A <= 8 && K <= 32) return rideshare_launch_geometry<8, 4, MODE>(p, io, B, s, entry_mask, batch_base);
  if (A <= 16 && K <= 32) return rideshare_launch_geometry<16, 2, MODE>(p, io, B, s, entry_mask, batch_base);
  if (A <= 16) return rideshare_launch_geometry<16, 4, MODE>(p, io, B, s, entry_mask, batch_base);
  if (K <= 32) return rideshare_launch_geometry<32, 1, MODE>(p, io, B, s, entry_mask, batch_base);
  return rideshare_launch_geometry<32, 2, MODE>(p, io, B, s, entry_mask, batch_base);
}

int rideshare_launch(const FrzRideshareParams* p, const FrzRideshareBuffers* io, int B, int mode, void* stream,
                     const uint8_t* entry_mask = nullptr, int batch_base = 0) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (mode == kRsStep) return rideshare_launch_mode<kRsStep>(p, io, B, s, entry_mask, batch_base);
  if (mode == kRsRefresh) return rideshare_launch_mode<kRsRefresh>(p, io, B, s, entry_mask, batch_base);
  return rideshare_launch_mode<kRsEntryRefresh>(p, io, B, s, entry_mask, batch_base);
}

}  // namespace
}  // namespace frz

extern "C" {

int frz_rideshare_step(const FrzRideshareParams* params, const FrzRideshareBuffers* io, int32_t parallel_envs,
                       void* stream) {
  const int status = frz::rideshare_validate(params, io, parallel_envs, "frz_rideshare_step");
  if (status != FRZ_OK) return status;
  if (io->actions == nullptr) {
    frz::set_error("frz_rideshare_step: actions is NULL");
    return FRZ_ERR_NULL;
  }
  return frz::rideshare_launch(params, io, parallel_envs, frz::kRsStep, stream);
}

int frz_rideshare_step_host(const FrzRideshareParams* params, const FrzRideshareBuffers* io, int32_t parallel_envs,
                            const FrzHostStep* host, void* stream) {
  const int status = frz::rideshare_validate(params, io, parallel_envs, "frz_rideshare_step_host");
  if (status != FRZ_OK) return status;
  if (io->actions == nullptr || io->rewards == nullptr) {
    frz::set_error("frz_rideshare_step_host: actions / rewards is NULL");
    return FRZ_ERR_NULL;
  }
  const size_t A = size_t(params->num_agents), K = size_t(params->capacity);
  const frz::HostArrays arrays{io->actions, io->rewards, io->terminated, io->truncated, io->control, params->num_agents,
                                 params, sizeof(FrzRideshareParams), io, sizeof(FrzRideshareBuffers)};
  return frz::run_host_pipeline(
      "frz_rideshare_step_host", host, arrays, parallel_envs, static_cast<cudaStream_t>(stream),
      [&](int first, int count, FrzControl* control, cudaStream_t slice_stream) {
        FrzRideshareParams p = *params;
        p.env_offset += first;  // schedule rows name global environment indices
        FrzRideshareBuffers slice = *io;
        const size_t e = size_t(first);
        slice.agents += e * A * 2;
        slice.passengers += e * K * FRZ_RS_PASSENGER_COLUMNS;
        slice.actions += e * A * 2;
        slice.rewards += e * A;
        slice.cumulative_rewards += e * A;
        slice.terminated += e;
        slice.truncated += e;
        slice.num_moves += e;
        slice.env_task_count += e;
        slice.agent_task_count += e * A;
        slice.task_mask += e * A * K;
        slice.self_obs += e * A * 4;
        slice.task_obs += e * K * FRZ_RS_TASK_COLUMNS;
        slice.control = control;
        return frz::rideshare_launch(&p, &slice, count, frz::kRsStep, slice_stream, nullptr, first);
      });
}

int frz_rideshare_refresh(const FrzRideshareParams* params, const FrzRideshareBuffers* io, int32_t parallel_envs,
                          void* stream) {
  const int status = frz::rideshare_validate(params, io, parallel_envs, "frz_rideshare_refresh");
  if (status != FRZ_OK) return status;
  return frz::rideshare_launch(params, io, parallel_envs, frz::kRsRefresh, stream);
}

int frz_rideshare_reset(const FrzRideshareParams* params, const FrzRideshareBuffers* io, int32_t parallel_envs,
                        const uint8_t* env_mask, void* stream) {
  int status = frz::rideshare_validate(params, io, parallel_envs, "frz_rideshare_reset");
  if (status != FRZ_OK) return status;
  if (io->init_agents == nullptr || io->init_passengers == nullptr || io->init_count == nullptr) {
    frz::set_error("frz_rideshare_reset: initial state is NULL");
    return FRZ_ERR_NULL;
  }
  const size_t total = size_t(parallel_envs) * params->capacity * FRZ_RS_PASSENGER_COLUMNS;
  const int grid = frz::persistent_grid(int((total + 255) / 256), 8);
  frz::rideshare_restore_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*params, *io, parallel_envs, env_mask);
  status = frz::check_launch("rideshare_restore_kernel");
  if (status != FRZ_OK) return status;
  // only the environments being reset admit their t = 0 passengers; the others are merely re-published
  return frz::rideshare_launch(params, io, parallel_envs, frz::kRsEntryRefresh, stream, env_mask);
}

int frz_rideshare_sample_actions(const FrzRideshareParams* params, const FrzRideshareBuffers* io, int32_t parallel_envs,
                                 uint64_t sampler_seed, void* stream) {
  const int status = frz::rideshare_validate(params, io, parallel_envs, "frz_rideshare_sample_actions");
  if (status != FRZ_OK) return status;
  const int total = parallel_envs * params->num_agents;
  const int grid = frz::persistent_grid((total + 255) / 256, 8);
  frz::rideshare_sample_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*params, *io, parallel_envs, sampler_seed);
  return frz::check_launch("rideshare_sample_kernel");
}

}  // extern "C"
