"""Tensor-core IPA vs fp64 reference and vs the SIMT kernel; timing at the bench shape (GPU box)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from debug_ipa_tc_common import make, head_major, split, ref, ops, dev, H

names = [("scalar", 0, 512), ("point", 512, 1280), ("pair", 1280, 1792), ("norm", 1792, 2048)]
for B, Lm, scale in ((3, 84, 1.5), (2, 56, 1.5), (130, 20, 1.5), (2, 200, 1.5), (2, 84, 100.0), (2, 256, 1.5), (2, 257, 1.5), (3, 300, 1.5), (2, 384, 1.5), (2, 512, 1.5), (1, 500, 100.0)):
    proj, rot, trans, pb, pv, hw, shape = make(B, Lm, seed=Lm, pos_scale=scale)
    r64 = ref(proj, rot, trans, pb, pv, hw, B, Lm)
    ws = ops.ipa_tc_workspace(shape, dev)
    pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
    sc_, pt_ = split(proj)
    for odt in (torch.float32, torch.bfloat16):
        o = ops.ipa_attention_tc_fwd(sc_, pt_, rot, trans, pbt, pvp, None, hw, shape, ws, out_dtype=odt)
        torch.cuda.synchronize()
        e = (o.double() - r64).abs()
        print(f"B={B} L={Lm} scale={scale} out={odt}:", {n: f"{e[:, a:b].max().item():.2e}/{r64[:, a:b].abs().max().item():.1f}" for n, a, b in names},
              "nan" if torch.isnan(o).any() else "")
# timing at the bench shape
B, Lm = 256, 84
proj, rot, trans, pb, pv, hw, shape = make(B, Lm)
ws = ops.ipa_tc_workspace(shape, dev); pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
out = torch.empty(B * Lm, 2048, dtype=torch.bfloat16, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
sc_, pt_ = split(proj)
print("tc   ms:", t(lambda: ops.ipa_attention_tc_fwd(sc_, pt_, rot, trans, pbt, pvp, None, hw, shape, ws, out=out)))
print("simt ms:", t(lambda: ops.ipa_attention_fwd(proj, rot, trans, pb, pv, None, hw, 1 / math.sqrt(48), shape, 1)))
