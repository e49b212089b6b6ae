"""Debug helper: run bench.mixed_leg's data through encode/decode and report which blocks differ."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import htscodecs_b200 as hb
from htscodecs_b200 import synth

n = 1 << 20
distinct = int(sys.argv[1]) if len(sys.argv) > 1 else 128
nblk = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
spec = synth.mixed_corpus(distinct, seed=5, size=n, ragged=False)
blocks = [synth.GENERATORS[g](b, m) for g, b, m, _, _ in spec]
orders1 = np.array([f | (hb.ORDER_RANS4x8 if meth else 0) for _, _, _, f, meth in spec], dtype=np.int64)
meth1 = np.array([meth for *_, meth in spec], dtype=np.uint8)
lib = hb.load_library()
cap = max(lib.hts_b200_compress_bound_4x8(n), max(hb.rans_compress_bound_4x16(n, int(f)) for f in set(orders1[meth1 == 0])))
cap = (cap + 15) // 16 * 16
ctx = hb.Context(0)
reps_t = (nblk + distinct - 1) // distinct
d_one = torch.from_numpy(np.concatenate(blocks)).cuda()
d_raw = d_one.repeat(reps_t)[: nblk * n].contiguous()
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
for it in range(3):
    comp_len.fill_(cap)
    ctx.compress_batch_dev(nblk, d_raw, raw_off, raw_len, d_comp, comp_off, comp_len, status, order)
    bad = torch.nonzero(status).flatten().tolist()
    print("iter", it, "encode bad status:", [(b, int(status[b]), spec[b % distinct]) for b in bad[:10]], len(bad))
    in_len = comp_len.clone()
    out_len.fill_(n)
    d_out.zero_()
    torch.cuda.synchronize()          # torch's stream is not the context's stream
    ctx.uncompress_batch_dev(nblk, d_comp, comp_off, in_len, d_out, raw_off, out_len, status, method)
    bad = torch.nonzero(status).flatten().tolist()
    print("iter", it, "decode bad status:", [(b, int(status[b]), spec[b % distinct]) for b in bad[:10]], len(bad))
    neq = (d_out.view(nblk, n) != d_raw.view(nblk, n)).any(dim=1)
    badb = torch.nonzero(neq).flatten().tolist()
    print("iter", it, "mismatching blocks:", len(badb))
    seen = set()
    for b in badb:
        k = b % distinct
        if k in seen: continue
        seen.add(k)
        diff = torch.nonzero(d_out.view(nblk, n)[b] != d_raw.view(nblk, n)[b]).flatten()
        print("   block", b, "spec", spec[k], "first diff", int(diff[0]), "ndiff", len(diff), "out_len", int(out_len[b]), "clen", int(in_len[b]))
        if len(seen) > 12: break
    # cross-check the first bad stream with the host API
    if badb:
        b = badb[0]; k = b % distinct
        c = bytes(d_comp[b * cap: b * cap + int(in_len[b])].cpu().numpy())
        o, st = ctx.uncompress_many([c], [n], [int(meth1[k])])
        print("   host-api decode of that stream ok:", st, o[0] == blocks[k].tobytes())
        c2, st2 = ctx.compress_many([blocks[k].tobytes()], [int(orders1[k])])
        print("   host-api encode equal:", c2[0] == c)
