#!/usr/bin/env python
"""SASS listings of the production kernels (cuobjdump on the in-tree objects; no GPU needed):
    python tools/dump_sass.py        -> profiles/sass/<kernel>.sass + profiles/sass/SUMMARY.txt
SUMMARY.txt holds, per kernel, the register count and the counts of the mnemonics that prove the
Blackwell path (UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM / STTM = tcgen05.ld / st, ...)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "videogpt_b200", "build")
OUT = os.path.join(ROOT, "profiles", "sass")
WANT = [  # (object, substring of the demangled name, short file name)
    ("gemm_pair_tcgen05.o", "gemm_bf16_pair_kernel<256, 0>", "gemm_pair_bn256_store"),
    ("gemm_pair_tcgen05.o", "gemm_bf16_pair_kernel<256, 2>", "gemm_pair_bn256_swiglu"),
    ("gemm_pair_tcgen05.o", "gemm_bf16_pair_kernel<192, 1>", "gemm_pair_bn192_residual"),
    ("attention_pair_tcgen05.o", "attn_pair_tcgen05_kernel<96, false>", "attn_pair_d96"),
    ("attention_pair_tcgen05.o", "attn_pair_tcgen05_kernel<128, false>", "attn_pair_d128"),
    ("elementwise.o", "rmsnorm_kernel", "rmsnorm"),
    ("elementwise.o", "rope_kv_append_kernel", "rope_kv_append"),
    ("elementwise.o", "embed_assemble_kernel", "embed_assemble"),
    ("elementwise.o", "linear_small_kernel<1>", "linear_small"),
    ("elementwise.o", "final_layer_kernel", "final_layer"),
    ("elementwise.o", "cfg_euler_kernel", "cfg_euler"),
    ("elementwise.o", "timestep_sinusoid_kernel", "timestep_sinusoid"),
    ("peer.o", "peer_barrier_kernel", "peer_barrier"),
]
KEY = ["UTCHMMA", "UTCBAR", "UTMALDG", "LDTM", "STTM", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "HMMA", "LDGSTS", "STL", "LDL",
       "LDG", "STG", "ATOM", "MEMBAR", "ELECT", "USETMAXREG", "UCGABAR"]
os.makedirs(OUT, exist_ok=True)
summary = []
for obj, sub, short in WANT:
    path = os.path.join(BUILD, obj)
    names = subprocess.run(["cuobjdump", "-elf", path], capture_output=True, text=True).stdout  # noqa
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    blocks = re.split(r"(?=\n\s*Function : )", sass)
    for b in blocks:
        m = re.search(r"Function : (\S+)", b)
        if not m:
            continue
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        if sub.replace(" ", "") not in dem.replace(" ", "").replace("(int)", ""):
            continue
        lines = [l.rstrip() for l in b.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        text = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in lines]
        with open(os.path.join(OUT, short + ".sass"), "w") as f:
            f.write(f"// {dem}\n// {obj}, sm_100a, cuobjdump -sass\n" + "\n".join(text) + "\n")
        ops = collections.Counter()
        for l in text:
            mm = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if mm:
                ops[mm.group(1)] += 1
        keys = {k: sum(v for o, v in ops.items() if o.startswith(k)) for k in KEY}
        res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        reg = "?"
        for blk in res.split("Function ")[1:]:
            if blk.startswith(m.group(1)):
                r = re.search(r"REG:(\d+)", blk)
                reg = r.group(1) if r else "?"
        summary.append(f"{short:28s} regs {reg:>4s}  instrs {len(text):5d}  " + " ".join(f"{k}={v}" for k, v in keys.items() if v))
        break
    else:
        summary.append(f"{short:28s} NOT FOUND ({sub})")
open(os.path.join(OUT, "SUMMARY.txt"), "w").write("\n".join(summary) + "\n")
print("\n".join(summary))
