"""One warm + a few launches of the tensor-core IPA at the bench shape (ncu target)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from debug_ipa_tc_common import make, head_major, split, ops, dev, H
B, Lm = int(os.environ.get("IPA_B", 256)), int(os.environ.get("IPA_L", 84))
proj, rot, trans, pb, pv, hw, shape = make(B, Lm)
ws = ops.ipa_tc_workspace(shape, dev); pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
out = torch.empty(B * Lm, 2048, dtype=torch.bfloat16, device=dev)
sc_, pt_ = split(proj)
both_ = torch.cat([sc_, pt_.to(torch.bfloat16)], dim=1)      # production layout: one projection writes scalar | bf16 point records
sc_, pt_ = both_[:, :sc_.shape[1]], both_[:, sc_.shape[1]:]
for _ in range(3):
    ops.ipa_attention_tc_fwd(sc_, pt_, rot, trans, pbt, pvp, None, hw, shape, ws, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
