"""Agreement of the reduced-precision evaluator paths with the fp32 path on real encoded positions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
from oracle import chess_shim  # noqa
from oracle.encoding_ref import encode_board, get_legal_actions
from conftest import random_playout_boards
from test_oracle_nn import load_case
from matrix0_b200.model import PolicyValueNet

g, cfg, sd = load_case(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden"), "r24")
N = int(os.environ.get("M0_PROBE_POSITIONS", "512"))
boards = random_playout_boards(max(20, N // 20), 120, seed=31)[::3][:N]
x = torch.from_numpy(np.stack([encode_board(b) for b in boards]))
legal = torch.from_numpy(np.stack([get_legal_actions(b) for b in boards])).cuda()
nets = {p: PolicyValueNet(cfg, device="cuda", precision=p) for p in ("fp32", "bf16", "fp16")}
for n in nets.values():
    n.load_state_dict(sd, strict=True)
pr, vr = nets["fp32"].forward(x)
neg = torch.full_like(pr, -1e30)
top_ref_legal = torch.where(legal, pr, neg).argmax(1)
for p in ("bf16", "fp16"):
    lg, v = nets[p].forward(x)
    all_top1 = (lg.argmax(1) == pr.argmax(1)).float().mean().item()
    legal_top1 = (torch.where(legal, lg, neg).argmax(1) == top_ref_legal).float().mean().item()
    print(p, "n", len(boards), "top1_all", round(all_top1, 4), "top1_legal", round(legal_top1, 4), "max|dv|", (v - vr).abs().max().item(),
          "max|dlogit|", (lg - pr).abs().max().item(), "logit_std", pr.std().item(), flush=True)
