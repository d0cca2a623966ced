"""Loops the tensor-core IPA operator for a few seconds while sampling nvidia-smi: is the kernel clock- / power-limited? (developer diagnostics)"""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from debug_ipa_tc_common import make, split, ops, dev, H
B, Lm = int(os.environ.get("IPA_B", 256)), int(os.environ.get("IPA_L", 84))
proj, rot, trans, pb, pv, hw, shape = make(B, Lm)
ws = ops.ipa_tc_workspace(shape, dev); pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
out = torch.empty(B * Lm, 2048, dtype=torch.bfloat16, device=dev)
sc_, pt_ = split(proj)
both_ = torch.cat([sc_, pt_.to(torch.bfloat16)], dim=1)
sc_, pt_ = both_[:, :sc_.shape[1]], both_[:, sc_.shape[1]:]
run = lambda: ops.ipa_attention_tc_fwd(sc_, pt_, rot, trans, pbt, pvp, None, hw, shape, ws, out=out)
samples, stop = [], False
def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True)
        samples.append(r.stdout.strip()); time.sleep(0.2)
for _ in range(5): run()
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(4):
    e0.record()
    for _ in range(4000): run()
    e1.record(); torch.cuda.synchronize()
    print(f"rep {rep}: {e0.elapsed_time(e1) / 4000 * 1e3:.1f} us per call")
stop = True; th.join()
print("\n".join(samples[::2]))
