"""One device-resident encode + decode of bench.mixed_leg's corpus (for an ncu launch list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import htscodecs_b200 as hb
from htscodecs_b200 import synth

n, distinct, nblk = 1 << 20, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
spec = synth.mixed_corpus(distinct, seed=5, size=n, ragged=False)
blocks = [synth.GENERATORS[g](b, m) for g, b, m, _, _ in spec]
orders1 = np.array([f | (hb.ORDER_RANS4x8 if meth else 0) for _, _, _, f, meth in spec], dtype=np.int64)
meth1 = np.array([meth for *_, meth in spec], dtype=np.uint8)
lib = hb.load_library()
cap = max(lib.hts_b200_compress_bound_4x8(n), max(hb.rans_compress_bound_4x16(n, int(f)) for f in set(orders1[meth1 == 0])))
cap = (cap + 15) // 16 * 16
ctx = hb.Context(0)
reps_t = (nblk + distinct - 1) // distinct
d_raw = torch.from_numpy(np.concatenate(blocks)).cuda().repeat(reps_t)[: nblk * n].contiguous()
order = torch.from_numpy(np.tile(orders1, reps_t)[:nblk].astype(np.int32)).cuda()
method = torch.from_numpy(np.tile(meth1, reps_t)[:nblk].copy()).cuda()
raw_off = torch.arange(nblk, dtype=torch.int64, device="cuda") * n
raw_len = torch.full((nblk,), n, dtype=torch.int32, device="cuda")
status = torch.zeros(nblk, dtype=torch.int32, device="cuda")
d_out = torch.zeros(nblk * n, dtype=torch.uint8, device="cuda")
d_comp = torch.empty(nblk * cap, dtype=torch.uint8, device="cuda")
comp_off = torch.arange(nblk, dtype=torch.int64, device="cuda") * cap
comp_len = torch.full((nblk,), cap, dtype=torch.int32, device="cuda")
out_len = torch.full((nblk,), n, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
import time
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    comp_len.fill_(cap); torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx.compress_batch_dev(nblk, d_raw, raw_off, raw_len, d_comp, comp_off, comp_len, status, order)
    t1 = time.perf_counter()
    in_len = comp_len.clone(); out_len.fill_(n); torch.cuda.synchronize(); t2 = time.perf_counter()
    ctx.uncompress_batch_dev(nblk, d_comp, comp_off, in_len, d_out, raw_off, out_len, status, method)
    t3 = time.perf_counter()
    print(f"iter {it}: enc {1e3*(t1-t0):.1f} ms  dec {1e3*(t3-t2):.1f} ms (host clock, synchronous calls)")
assert int((status != 0).sum()) == 0 and torch.equal(d_out, d_raw)
