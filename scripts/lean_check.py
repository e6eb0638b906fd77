# coding: utf-8
"""Lean stage path against the generic stage path on the same inputs (free running, same seed; the lean path adds the
x-part of a gate pre-activation before the y-part, so the two differ by fp32 rounding):  python scripts/lean_check.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from scripts.sweep import CFGS  # noqa: E402
from wavenet_vocoder_b200 import WaveNet  # noqa: E402

ok = True
for name in ("cfg2", "cfg5"):
    kw = CFGS[name]
    torch.manual_seed(0)
    m = WaveNet(**kw).eval()
    with torch.no_grad():
        O = kw["out_channels"]
        m.last_conv_layers[3].bias[2 * (O // 3):] -= 3.0
    m = m.cuda()
    eng = m._get_engine()
    T = 700
    c = torch.randn(1, T, kw["cin_channels"], device="cuda")
    outs = {}
    for tag, env in (("generic", {"WN_LEAN": "0"}), ("lean", {"WN_LEAN": "1"}),
                     ("lean_fast_gate", {"WN_LEAN": "1", "WN_FAST_GATE": "1"})):
        for k in ("WN_LEAN", "WN_FAST_GATE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        y = eng.generate(B=1, T=T, c=c, seed=5)[0]
        torch.cuda.synchronize()
        outs[tag] = y.clone()
    for k in ("WN_LEAN", "WN_FAST_GATE"):
        os.environ.pop(k, None)
    d = float(((outs["generic"] - outs["lean"]) ** 2).mean().sqrt())
    same = d <= 1e-4
    rms = float(((outs["generic"] - outs["lean_fast_gate"]) ** 2).mean().sqrt())
    first = int((outs["generic"] != outs["lean_fast_gate"]).float().argmax()) if rms > 0 else -1
    print(name, "lean vs generic RMS %.3g:" % d, same, "| fast gate: RMS diff %.3g, first differing sample %d" % (rms, first),
          "| finite:", bool(torch.isfinite(outs["lean"]).all()))
    ok = ok and same
print("LEAN_CHECK", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)
