import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import htscodecs_b200 as hb
from htscodecs_b200 import synth
import bench
ctx = hb.Context(0)
n = 1 << 20
for name, gen, f in (("pack_acgt", "acgt", 0x80), ("pack_o1_acgt", "acgt", 0x81), ("stripe4_u32", "u32", 0x408), ("stripe4_o1_u32", "u32", 0x409)):
    blocks = [synth.GENERATORS[gen](i, n) for i in range(16)]
    try:
        print(name, bench.path_sweep(ctx, torch, hb, blocks, 4096, reps=1, legs=None, one=(name, f)), "scratch", ctx.scratch_bytes >> 20, "MiB", flush=True)
    except Exception as e:
        print(name, "FAILED", repr(e)[:300], "scratch", ctx.scratch_bytes >> 20, "MiB", flush=True)
