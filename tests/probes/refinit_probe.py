"""Reduced-precision evaluator paths vs the unmodified reference's fp32 outputs on the REFERENCE'S OWN random init
(tests/golden/refinit_golden.npz): top-1 agreement, value error, and the logit error against this repo's fp32 CUDA path.
    python tests/probes/refinit_probe.py [fp16 bf16 ...]  ->  one JSON line per precision"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from oracle import chess_shim  # noqa
import chess
from oracle.encoding_ref import encode_board, get_legal_actions
from matrix0_b200.model import NetConfig, PolicyValueNet

g = np.load(os.path.join(ROOT, "tests", "golden", "refinit_golden.npz"))
d = json.loads(str(g["cfg"]))
cfg = NetConfig(**{k: v for k, v in d.items() if k in set(NetConfig.__dataclass_fields__)})
N = int(os.environ.get("M0_PROBE_POSITIONS", "2304"))
boards = [chess.Board(str(f)) for f in g["fens"][:N]]
x = torch.from_numpy(np.stack([encode_board(b) for b in boards]))
legal = torch.from_numpy(np.stack([get_legal_actions(b) for b in boards])).cuda()


def run(precision):
    net = PolicyValueNet(cfg, device="cuda", precision=precision, seed=0)
    ps, vs = [], []
    for i in range(0, N, 1152):
        p, v = net.forward(x[i:i + 1152])
        ps.append(p.float())
        vs.append(v.float())
    return torch.cat(ps), torch.cat(vs)


pr, vr = run("fp32")
neg = torch.full_like(pr, -1e30)
for prec in (sys.argv[1:] or ["fp32", "fp16", "bf16"]):
    p, v = (pr, vr) if prec == "fp32" else run(prec)
    top1 = p.argmax(1).cpu().numpy()
    top1l = torch.where(legal, p, neg).argmax(1).cpu().numpy()
    dl = (p - pr).abs()
    print(json.dumps({"precision": prec, "n": N, "env": {k: v for k, v in os.environ.items() if k.startswith("M0_")},
                      "top1_vs_reference_fp32": float((top1 == g["top1"][:N]).mean()),
                      "top1_legal_vs_reference_fp32": float((top1l == g["top1_legal"][:N]).mean()),
                      "max_dv_vs_reference_fp32": float(np.abs(v.cpu().numpy() - g["value"][:N]).max()),
                      "max_dlogit_vs_own_fp32": float(dl.max()), "mean_dlogit_vs_own_fp32": float(dl.mean()),
                      "rms_dlogit_centered": float(((p - p.mean(1, keepdim=True)) - (pr - pr.mean(1, keepdim=True))).pow(2).mean().sqrt()),
                      "logit_std": float(pr.std(1).mean())}), flush=True)
