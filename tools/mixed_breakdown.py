"""Per-category device-resident encode/decode time of the mixed corpus (bench.mixed_leg's data)."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import htscodecs_b200 as hb
from htscodecs_b200 import synth

n = 1 << 20
distinct, nblk = 128, 4096
spec = synth.mixed_corpus(distinct, seed=5, size=n, ragged=False)
cats = collections.OrderedDict()
for k, (g, b, m, f, meth) in enumerate(spec):
    cats.setdefault((g, f, meth), []).append(k)
ctx = hb.Context(0)
lib = hb.load_library()
stream = torch.cuda.ExternalStream(ctx.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
only = sys.argv[1:]
for (g, f, meth), ks in cats.items():
    cnt = len(ks) * nblk // distinct
    name = f"{g}/{f:#x}/{'4x8' if meth else '4x16'}"
    if only and not any(o in name for o in only):
        continue
    blocks = [synth.GENERATORS[g](spec[k][1], n) for k in ks[:8]]
    order_v = f | (hb.ORDER_RANS4x8 if meth else 0)
    cap = lib.hts_b200_compress_bound_4x8(n) if meth else hb.rans_compress_bound_4x16(n, f)
    cap = (cap + 15) // 16 * 16
    d_one = torch.from_numpy(np.concatenate(blocks)).cuda()
    d_raw = d_one.repeat((cnt + len(blocks) - 1) // len(blocks))[: cnt * n].contiguous()
    order = torch.full((cnt,), order_v, dtype=torch.int32, device="cuda")
    method = torch.full((cnt,), meth, dtype=torch.uint8, device="cuda")
    raw_off = torch.arange(cnt, dtype=torch.int64, device="cuda") * n
    raw_len = torch.full((cnt,), n, dtype=torch.int32, device="cuda")
    status = torch.zeros(cnt, dtype=torch.int32, device="cuda")
    d_out = torch.empty(cnt * n, dtype=torch.uint8, device="cuda")
    d_comp = torch.empty(cnt * cap, dtype=torch.uint8, device="cuda")
    comp_off = torch.arange(cnt, dtype=torch.int64, device="cuda") * cap
    comp_len = torch.full((cnt,), cap, dtype=torch.int32, device="cuda")
    out_len = torch.full((cnt,), n, dtype=torch.int32, device="cuda")
    te, td = [], []
    for it in range(3):
        comp_len.fill_(cap); torch.cuda.synchronize()
        e0.record(stream)
        ctx.compress_batch_dev(cnt, d_raw, raw_off, raw_len, d_comp, comp_off, comp_len, status, order, sync=(it == 0))
        e1.record(stream); torch.cuda.synchronize(); te.append(e0.elapsed_time(e1))
        assert int((status != 0).sum()) == 0
        in_len = comp_len.clone()
        out_len.fill_(n); torch.cuda.synchronize()
        e0.record(stream)
        ctx.uncompress_batch_dev(cnt, d_comp, comp_off, in_len, d_out, raw_off, out_len, status, method, sync=(it == 0))
        e1.record(stream); torch.cuda.synchronize(); td.append(e0.elapsed_time(e1))
        assert int((status != 0).sum()) == 0 and torch.equal(d_out, d_raw)
    ratio = float(in_len.to(torch.int64).sum()) / (cnt * n)
    print(f"{name:22s} blocks {cnt:5d}  enc {min(te[1:]):8.2f} ms  dec {min(td[1:]):8.2f} ms  ratio {ratio:.3f}  first byte {int(d_comp[0]):#x}", flush=True)
    del d_raw, d_out, d_comp
