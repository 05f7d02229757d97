"""Independent host-side Philox4x32-10 + the engine's counter scheme -- TEST INFRASTRUCTURE.

The production step kernels (``wildfire_step_kernel<..., INJECTED=false>``, ``cyber_step_tiled_kernel<false, ...>``)
draw their randomness in-kernel.  To pin those instantiations to the oracle bit-for-bit, this module restates

1. Philox4x32-10 (Salmon et al., SC'11; the same round function and Weyl key schedule as csrc/frz_common.cuh:33-54),
   vectorised over numpy uint32 arrays and checked against the published known-answer vectors
   (tests/test_philox_ref.py), and
2. which counter / word feeds which event of which cell / agent / node (csrc/frz_wildfire.cu "randomness" section,
   csrc/frz_cyber.cu:115-174) as a plain table lookup,

and expands them into uniforms shaped like the reference's ``generator.generate`` output -- wildfire ``(3, B, H, W)`` +
``(5, B, A)`` (wildfire.py:409-410, 488-532), cybersecurity ``(1, B, N)`` + ``(1, B, n)`` -- which the oracle
consumes.  The GPU parity tests then compare Philox-mode rollouts with the oracle on those uniforms.

Counter of every call: ``(env_lo, step_lo, stream, step_hi ^ env_hi)`` with env = GLOBAL environment index
(``env_offset`` + local), step = the control block's step counter (0 for the first step after a reset), key = the
64-bit seed of ``reset(seed=...)``.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(key, counter):
    """key = (k0, k1) python ints; counter = 4 broadcastable uint32 arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*[np.asarray(c, dtype=np.uint32) for c in counter])
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0.astype(np.uint64)
        p1 = M1 * c2.astype(np.uint64)
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint32(k0), lo1, hi0 ^ c3 ^ np.uint32(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u01(bits):
    """24 random bits -> fp32 uniform on the grid k * 2^-24 (csrc/frz_common.cuh:57)."""
    return ((np.asarray(bits, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def _calls(seed: int, envs: np.ndarray, step: int, streams: np.ndarray) -> np.ndarray:
    """uint32 [len(envs), len(streams), 4]: the four words of Philox call ``stream`` of every environment."""
    envs = np.asarray(envs, dtype=np.uint64)
    env_lo = (envs & MASK32).astype(np.uint32)[:, None]
    env_hi = (envs >> np.uint64(32)).astype(np.uint32)[:, None]
    step_lo = np.uint32(step & 0xFFFFFFFF)
    step_hi = np.uint32((step >> 32) & 0xFFFFFFFF) ^ env_hi
    words = philox4x32_10((seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF),
                          (env_lo, step_lo, np.asarray(streams, dtype=np.uint32)[None, :], step_hi))
    return np.stack(words, axis=2)


# ------------------------------------------------------------------------------------------------ wildfire


def wildfire_geometry(H: int, W: int, A: int):
    """(lanes per environment, cells per lane) the dispatcher picks (csrc/frz_wildfire.cu pick_geometry)."""
    HW = H * W
    if HW <= 8 and A <= 8:
        return 8, 1
    if HW <= 32 and A <= 8:
        return 8, (HW + 7) // 8
    if HW <= 16 and A <= 16:
        return 16, 1
    if HW <= 128 and A <= 16:
        return 16, (HW + 15) // 16
    if HW <= 32:
        return 32, 1
    if HW <= 64:
        return 32, 2
    if HW <= 128:
        return 32, 4
    return 32, 8


def wildfire_layout(H: int, W: int, A: int):
    """The word layout of the production kernel for this grid: dict with the geometry and, per cell / agent event, the
    (stream, word) pair that feeds it.  Streams are the third Philox counter word."""
    G, CPL = wildfire_geometry(H, W, A)
    HW = H * W
    calls = (2 * CPL + 3) // 4
    cells_in_smem = G == 16 and CPL > 4
    split = cells_in_smem and 2 * ((CPL + 3) // 4) == calls
    calls_a = calls // 2 if split else calls
    grow_word = (lambda i: i) if split else (lambda i: 2 * i)
    spread_word = (lambda i: 4 * calls_a + i) if split else (lambda i: 2 * i + 1)

    def lane_word(sub, word):  # word `word` of lane `sub` -> (stream, word within the call)
        return (word // 4) * G + sub, word % 4

    grow = np.zeros((HW, 2), np.int64)
    spread = np.zeros((HW, 2), np.int64)
    for c in range(HW):
        i, sub = divmod(c, G)
        grow[c] = lane_word(sub, grow_word(i))
        spread[c] = lane_word(sub, spread_word(i))

    spare_own = 4 * calls - 2 * CPL
    feed = False
    if CPL > 1:
        last_row_cells = HW - G * (CPL - 1)
        spare_lanes = G - max(last_row_cells, 0)
        feed = (A if spare_own == 2 else 2 * A) <= spare_lanes
    agent = np.zeros((A, 4, 2), np.int64)  # the four words of every agent
    for a in range(A):
        if feed:
            first, second = (G - 1 - a) & (G - 1), (G - 1 - A - a) & (G - 1)
            spare_a, spare_b = lane_word(first, grow_word(CPL - 1)), lane_word(first, spread_word(CPL - 1))
            if spare_own == 2:
                own_a = CPL if split else 2 * CPL
                own_b = 4 * calls_a + CPL if split else 2 * CPL + 1
                agent[a] = (lane_word(a, own_a), lane_word(a, own_b), spare_a, spare_b)
            else:
                agent[a] = (spare_a, spare_b, lane_word(second, grow_word(CPL - 1)),
                            lane_word(second, spread_word(CPL - 1)))
        else:
            agent[a] = [(0x80000000 | a, j) for j in range(4)]
    return dict(G=G, CPL=CPL, calls=calls, split=split, cells_in_smem=cells_in_smem, spare_lanes_feed_agents=feed,
                grow=grow, spread=spread, agent=agent)


def wildfire_uniforms(seed: int, step: int, envs, H: int, W: int, A: int):
    """(u_field f32 [3, B, H, W], u_agent f32 [5, B, A]) the production kernel uses at ``step`` for the global
    environment indices ``envs``.  Events: field 0 increase, 1 decrease (one shared word: a cell consumes one or the
    other), 2 spread; agent 0 suppressant decrease and 2 refill (shared word: fight vs refill action), 1 equipment,
    3 capacity pick, 4 tank switch."""
    layout = wildfire_layout(H, W, A)
    envs = np.asarray(envs)
    B, HW = len(envs), H * W
    streams = sorted({int(s) for table in (layout['grow'], layout['spread'], layout['agent'].reshape(-1, 2))
                      for s in table[:, 0]})
    index = {s: i for i, s in enumerate(streams)}
    words = _calls(seed, envs, step, np.asarray(streams, dtype=np.uint64).astype(np.uint32))  # [B, S, 4]

    def gather(table):  # [..., 2] (stream, word) -> uint32 [B, ...]
        flat = table.reshape(-1, 2)
        columns = np.asarray([index[int(s)] for s in flat[:, 0]])
        return words[:, columns, flat[:, 1]].reshape((B, ) + table.shape[:-1])

    grow, spread = u01(gather(layout['grow'])), u01(gather(layout['spread']))
    u_field = np.stack([grow, grow, spread], axis=0).reshape(3, B, H, W)
    agent = u01(gather(layout['agent']))  # [B, A, 4]
    u_agent = np.stack([agent[..., 0], agent[..., 1], agent[..., 0], agent[..., 2], agent[..., 3]], axis=0)
    return np.ascontiguousarray(u_field), np.ascontiguousarray(u_agent)


# ------------------------------------------------------------------------------------------------ cybersecurity


def cyber_uniforms(seed: int, step: int, envs, num_nodes: int, num_agents: int):
    """(u_network f32 [1, B, N], u_agent f32 [1, B, n]): node j = word j % 4 of stream j // 4; agent a (attackers
    first) = word a % 4 of stream 0x80000000 | a // 4 (csrc/frz_cyber.cu:147, 171)."""
    envs = np.asarray(envs)
    node_calls = _calls(seed, envs, step, np.arange((num_nodes + 3) // 4, dtype=np.uint32))
    agent_calls = _calls(seed, envs, step, np.uint32(0x80000000) | np.arange((num_agents + 3) // 4, dtype=np.uint32))
    u_network = u01(node_calls.reshape(len(envs), -1)[:, :num_nodes])
    u_agent = u01(agent_calls.reshape(len(envs), -1)[:, :num_agents])
    return u_network[None], u_agent[None]


# ------------------------------------------------------------------------------------------------ action sampler


def sampled_actions(sampler_seed: int, step: int, envs, counts: np.ndarray):
    """The wildfire / rideshare-style uniform legal-action sampler (csrc/frz_wildfire.cu wildfire_sample_kernel):
    agent a of env e draws word 0 of stream 0xC0000000 | a; k = min(int(u * (n + 1)), n); k == n is the no-op."""
    envs = np.asarray(envs)
    A = counts.shape[1]
    words = _calls(sampler_seed, envs, step, np.uint32(0xC0000000) | np.arange(A, dtype=np.uint32))[:, :, 0]
    u = u01(words)
    k = np.minimum((u * (counts + 1).astype(np.float32)).astype(np.float32).astype(np.int64), counts)
    return np.stack([k, np.where(k == counts, -1, 0)], axis=2).astype(np.int32)
