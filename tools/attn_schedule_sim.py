#!/usr/bin/env python
"""Event simulation of attn_pair_tcgen05_kernel's software pipeline (early-S path, two query tiles):
an in-order tensor pipe fed by ONE issuing warp that walks the fixed order
    [wait s_free(A) -> S_A(j+1)] [wait p_full(A) -> PV_A(j)] [wait s_free(B) -> S_B(j+1)] [wait p_full(B) -> PV_B(j)]
with blocking waits, two softmax warpgroups of constant latency, and the shared P buffer
(P_A(j) after PV_B(j-1), P_B(j) after PV_A(j)).  Prints the steady-state period per pair of tiles
and KV tile for a grid of softmax latencies and hand-over latencies (cycles).  No GPU needed; the
durations are the measured issue costs (profiles/r01h_umma_rate.txt): S = 6 x 107, P V = 8 x 58."""


def period(t_softmax, handover, s=642, pv=464, n=80, t_pwrite=100):
    pipe_free, t = 0.0, 0.0
    s_done, pv_done, p_ready = {}, {}, {}

    def issue(dur):
        nonlocal pipe_free, t
        pipe_free = max(pipe_free, t) + dur
        return pipe_free

    s_done[(0, 0)] = issue(s)
    s_done[(1, 0)] = issue(s)

    def softmax(x, j):
        start = s_done[(x, j)] + handover
        done = start + t_softmax
        need = (0, j) if x == 1 else (1, j - 1)          # previous reader of the shared P buffer
        if need[1] >= 0:
            done = max(done, pv_done.get(need, 0.0) + handover)
        p_ready[(x, j)] = done + t_pwrite
        return start + 60                                  # S in registers: s_free

    for j in range(n):
        for x in (0, 1):
            if j + 1 < n:
                t = max(t, softmax(x, j) + handover)
                s_done[(x, j + 1)] = issue(s)
            softmax(x, j)
            t = max(t, p_ready[(x, j)] + handover)
            pv_done[(x, j)] = issue(pv)
    return (pv_done[(1, n - 1)] - pv_done[(1, n // 2)]) / (n - 1 - n // 2)


def period_single_tile(t_softmax, handover, s=444, pv=464, n=80, t_pwrite=100, p_bufs=1):
    """Candidate design: ONE query tile per CTA, S double-buffered (S(j+2) issued into the buffer softmax(j)
    has just read), S in TS form with Q in tensor memory (6 x 74 cycles; TMEM: 2 x 128 S + 96 O + 64 P +
    48 Q = 464 columns), one P buffer.  Same hand-over model as above; period per KV tile."""
    pipe_free, t = 0.0, 0.0
    s_done, pv_done, sm_done = {}, {}, {-1: 0.0}

    def issue(dur):
        nonlocal pipe_free, t
        pipe_free = max(pipe_free, t) + dur
        return pipe_free

    s_done[0] = issue(s)
    s_done[1] = issue(s)
    for j in range(n):
        done = max(s_done[j] + handover, sm_done[j - 1]) + t_softmax
        if j - p_bufs >= 0:
            done = max(done, pv_done[j - p_bufs] + handover)
        sm_done[j] = done + t_pwrite
        t = max(t, sm_done[j] + handover)
        pv_done[j] = issue(pv)
        if j + 2 < n:
            s_done[j + 2] = issue(s)
    return (pv_done[n - 1] - pv_done[n // 2]) / (n - 1 - n // 2)


if __name__ == "__main__":
    lat = (0, 1100, 1600, 2300, 3000, 4000)
    print("period per pair of tiles and KV tile (tensor-pipe work: 2212 cycles)")
    print("hand-over \\ softmax latency " + "".join(f"{x:>7d}" for x in lat))
    for h in (100, 200, 300, 400, 600, 800):
        print(f"{h:>26d} " + "".join(f"{period(x, h):7.0f}" for x in lat))
    print("measured (r01h, old grid): 83 us per launch = 4400-5200 cycles per period; tensor-pipe chain alone 53 us = 2800-3300")
    print("\ncandidate: one query tile per CTA, S double-buffered, S in TS form -- period per KV tile for ONE tile")
    print("(compare with HALF the numbers above; tensor-pipe work 908, MUFU 1024 per tile.  Reality check: the first")
    print(" tcgen05 kernel, attention_tcgen05.cu, IS one tile with double-buffered S and P (SS form) and measured")
    print(" 101 us = 3000-3800 cycles per tile -- a real softmax pass is much longer than the latencies assumed here)")
    lat1 = (700, 1100, 1300, 1600)
    print("hand-over \\ softmax latency " + "".join(f"{x:>7d}" for x in lat1))
    for h in (100, 300, 500, 700):
        print(f"{h:>26d} " + "".join(f"{period_single_tile(x, h):7.0f}" for x in lat1))
