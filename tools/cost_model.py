#!/usr/bin/env python
"""Analytic cost model of one Euler step (runs anywhere, no GPU): tiles, waves and tensor-pipe issue
cycles of the four projection GEMMs and of the attention launch, from the per-instruction costs
MEASURED on B200 (profiles/r01h_umma_rate.txt: a tcgen05.mma with both operands in shared memory
costs N/2 + 43 cycles per K = 16 step, N/2 + 10 with A in tensor memory) and the kernels' actual
tiling rules (gemm_tcgen05.cu::pick_pair_block_n, attention_pair_tcgen05.cu grid order).

    python tools/cost_model.py [--mhz 1600]

Prints, per workload / batch size: predicted microseconds per launch next to the measured ones
where profiles/ has them, and what the queued changes (skinny tail kernel, batching) are worth.
It is a planning tool: the numbers that count are the measured ones in profiles/.
"""
import argparse
import math

H, I, HEADS, D, LAYERS = 3072, 8192, 32, 96, 32
CLUSTERS = 74                      # 148 SMs / 2
SS_OVERHEAD, TS_OVERHEAD = 43, 10  # cycles per tcgen05.mma beyond N/2 (measured, r01h_umma_rate.txt)
MEASURED_US = {("cfg2", "qkv"): 92, ("cfg2", "o"): 38, ("cfg2", "gate_up"): 153, ("cfg2", "down"): 94,
               ("cfg2", "attention"): 83}     # profiles/r01c_gemm_sweep_pair.txt, r01h (burst clocks, kernel alone)


def waves(rows, n, bn):
    return math.ceil(math.ceil(rows / 256) * math.ceil(n / bn) / CLUSTERS)


def pick_block_n(rows, n):
    """gemm_tcgen05.cu::pick_pair_block_n: minimise waves x width / fill efficiency."""
    best = None
    for bn, eff in ((256, 1.0), (192, 0.88)):
        if n % bn:
            continue
        cost = waves(rows, n, bn) * bn / eff
        if best is None or cost < best[0]:
            best = (cost, bn)
    return best[1]


def gemm_cycles(rows, n, k, skinny=False):
    """Persistent CTA-pair kernel: every cluster walks ceil(tiles / 74) tiles of K/16 MMAs each."""
    tail = rows % 256
    extra = 0
    if skinny and 0 < tail <= 32 and rows > 256 and n % 256 == 0:
        full, main = pick_block_n(rows, n), pick_block_n(rows - tail, n)
        if waves(rows - tail, n, main) * main / (1.0 if main == 256 else 0.88) + 128 < \
                waves(rows, n, full) * full / (1.0 if full == 256 else 0.88):
            # tail kernel: one 256-column tile per cluster, N = 16 MMAs, bound by streaming W (64 B/clk/SM)
            extra = math.ceil(n / 256 / CLUSTERS) * (k // 64) * (128 * 64 * 2 + 1024) / 64
            rows -= tail
    bn = pick_block_n(rows, n)
    return waves(rows, n, bn) * (k // 16) * (bn / 2 + SS_OVERHEAD) + extra, bn


def attention_cycles(seqs, head_fastest=True):
    """seqs: [(q_rows, kv_len)] per sequence.  Unit = 128 x 128 tile: 6 S MMAs (SS) + 8 P V MMAs (TS)
    at head_dim 96.  CTAs = pairs of query tiles, one per SM at a time, list-scheduled in grid order:
    (head, query pair, sequence) with the head innermost (current), or the query pair innermost (the
    grid of the r01h measurements)."""
    unit = (D // 16) * (128 / 2 + SS_OVERHEAD) + (128 // 16) * (D / 2 + TS_OVERHEAD)
    q_pairs = max(math.ceil(q / 256) for q, _ in seqs)
    ctas = []
    for q, kv in seqs:
        work = [min(2, max(0, math.ceil((q - p * 256) / 128))) * math.ceil(kv / 128) for p in range(q_pairs)]
        order = [w for w in work for _ in range(HEADS)] if head_fastest else [w for _ in range(HEADS) for w in work]
        ctas += [w for w in order if w]
    sms = [0.0] * 148
    for c in ctas:                                   # hardware dispatch: next CTA to the first free SM
        i = min(range(148), key=sms.__getitem__)
        sms[i] += c
    useful = sum(q * kv for q, kv in seqs) * HEADS / (128 * 128)
    return max(sms) * unit, sum(ctas) / 148 * unit, useful / 148 * unit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mhz", type=float, default=1600.0, help="SM clock under load (bench.py clocks.sm_mhz)")
    args = ap.parse_args()
    us = lambda cyc: cyc / args.mhz
    geoms = {"cfg2": (4, 4, 258, 1), "cfg2 x4 videos (cfg4, --batch 4)": (4, 4, 258, 4),
             "cfg2 x8 videos (cfg4, --batch 8)": (4, 4, 258, 8), "cfg3": (32, 4, 258, 1), "cfg5": (4, 4, 1026, 1)}
    shapes = (("qkv", 3 * H, H), ("o", H, H), ("gate_up", 2 * I, H), ("down", H, I))
    print(f"clock {args.mhz:.0f} MHz; SS mma = N/2 + {SS_OVERHEAD} cycles, TS = N/2 + {TS_OVERHEAD}; {CLUSTERS} CTA pairs\n")
    for name, (n_ctx, n_gen, bl, vids) in geoms.items():
        t_gen, t_ctx = n_gen * bl, n_ctx * bl
        rows = 2 * t_gen * vids
        print(f"{name}: M = {rows} rows per GEMM")
        tot = {False: 0.0, True: 0.0}
        for sk in (False, True):
            for nm, n, k in shapes:
                cyc, bn = gemm_cycles(rows, n, k, skinny=sk)
                tot[sk] += cyc
                if not sk:
                    meas = MEASURED_US.get((name, nm))
                    ideal = rows * n * k / (256 * 16 * 128 / 64) / CLUSTERS      # 256 x N x 16 per N/2 cycles
                    print(f"   {nm:8s} BN {bn}: {us(cyc):7.1f} us  (ideal pipe {us(ideal):6.1f} us"
                          + (f", measured {meas} us at burst clocks)" if meas else ")"))
        seqs = [(t_gen, t_ctx + t_gen)] * vids + [(t_gen, t_gen)] * vids
        mk, even, useful = attention_cycles(seqs)
        mk_old = attention_cycles(seqs, head_fastest=False)[0]
        meas = MEASURED_US.get((name, "attention"))
        print(f"   attention : makespan {us(mk):6.1f} us (query pair innermost, the r01h grid: {us(mk_old):6.1f}), evenly spread "
              f"{us(even):6.1f}, without padding {us(useful):6.1f} (tensor-pipe chain only"
              + (f"; measured {meas} us with softmax, 53 us chain alone, both on the r01h grid)" if meas else ")"))
        step = LAYERS * (tot[False] + mk)
        print(f"   GEMMs per layer {us(tot[False]):6.1f} us -> with the skinny tail kernel {us(tot[True]):6.1f} us "
              f"({100 * (tot[True] / tot[False] - 1):+.1f} %)")
        print(f"   per row: {us(tot[False]) / rows * 1e3:6.2f} ns GEMM/layer;  tensor-pipe floor of a step {us(step) / 1e3:6.2f} ms\n")


if __name__ == "__main__":
    main()
