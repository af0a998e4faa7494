"""Diagnostic (GPU): per-tensor relative L2 errors of the fused stage against the float64 oracle on the same device.

  python tests/tools/diag_bf16_error.py [B C H L [A]] [--summary]        default 12 512 8 8 8

Columns: dsfuse fp32 mode, dsfuse bf16 mode (default build), dsfuse bf16 with fp32 dL/d(LayerNorm out) (DSF_LN_DY_BF16=0), stock
torch.autocast(bf16) running the oracle code, and dsfuse bf16 against the float64 oracle evaluated WITH dsfuse's ReLU decisions
("matched", see tests/parity_util.py).  --summary prints the worst tensor per class instead of every tensor."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import parity_util as U  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
summary = "--summary" in sys.argv
B, C, H, L = [int(v) for v in (args[:4] if len(args) >= 4 else (12, 512, 8, 8))]
A = int(args[4]) if len(args) >= 5 else 8
dev = torch.device("cuda")
pb = U.make_problem(B, C, H, L, A, dev)
ref = U.run_oracle(pb)
cols = {}
cols["f32"] = U.errors(U.run_dsfuse(pb, torch.float32), ref, L)
cap = {}
mine = U.run_dsfuse(pb, torch.bfloat16, capture=cap)
cols["bf16"] = U.errors(mine, ref, L)
os.environ["DSF_LN_DY_BF16"] = "0"
cols["bf16_lnfp32"] = U.errors(U.run_dsfuse(pb, torch.bfloat16), ref, L)
os.environ.pop("DSF_LN_DY_BF16")
cols["autocast"] = U.errors(U.run_oracle(pb, torch.float32, autocast=True), ref, L)
flip = U.flip_fraction(cap, U.oracle_relu_decisions(pb))
ref_m = U.run_oracle(pb, relu_masks=cap)
cols["bf16_matched"] = U.errors(mine, ref_m, L)
names = list(cols)
print("B=%d C=%d H=%d L=%d A=%d T=%d   relative L2 error vs the float64 oracle" % (B, C, H, L, A, pb["T"]))
print("ReLU decisions of the bf16 run that differ from float64: %.3e of the active units (worst block); sqrt = %.2e" % (flip, flip ** 0.5))
print("%-36s " % "tensor" + " ".join("%12s" % n for n in names))
keys = list(cols["bf16"])
if summary:
    classes = ["out", "gin", "g/pos_emb", "ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias", "attn.key.weight", "attn.key.bias", "attn.query.weight",
               "attn.query.bias", "attn.value.weight", "attn.value.bias", "attn.proj.weight", "attn.proj.bias", "mlp.0.weight", "mlp.0.bias",
               "mlp.2.weight", "mlp.2.bias", "ln_f.weight", "ln_f.bias"]
    for c in classes:
        ks = [k for k in keys if (k.startswith(c) if c in ("out", "gin", "g/pos_emb") else k.endswith(c))]
        print("%-36s " % (c + " (worst of %d)" % len(ks)) + " ".join("%12.2e" % max(cols[n][k] for k in ks) for n in names))
else:
    for k in keys:
        print("%-36s " % k + " ".join("%12.2e" % cols[n][k] for n in names))
print("WORST " + "  ".join("%s: %.2e %s" % ((n,) + U.worst({k: v for k, v in cols[n].items() if not k.endswith("key.bias")})) for n in names))
