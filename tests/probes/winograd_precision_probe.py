"""Would a Winograd F(2x2, 3x3) tower survive the precision gate?  CPU emulation, no kernel involved ("the lever
that follows from the energy bound").

The 48 3x3 convolutions of the residual tower are evaluated by the functional fp32 forward of oracle/nn_ref.py with the convolution
replaced by an emulation of what a 16-bit tensor-core kernel would compute, everything else staying fp32:
  direct   : activations and weights rounded to the operand format, fp32 accumulation (today's conv_pair_kernel, convolutions only)
  winograd : V = B^T d B on the rounded activations (fp32 adds, then rounded to the operand format: it is a tensor-core operand),
             U = G g G^T in fp32 rounded to the operand format, 16 products with fp32 accumulation, Y = A^T M A in fp32
Both are scored like tests/test_nn_gpu.py::test_refinit_fp16_gate: top-1 agreement with the unmodified reference's fp32 outputs on the
reference's own random init (tests/golden/refinit_golden.npz), all logits and legal moves only, and max |delta value|.

    python tests/probes/winograd_precision_probe.py [--positions 2304] [--formats fp16,bf16]  ->  one JSON line per (format, mode)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.nn.functional as F
from oracle import chess_shim  # noqa: F401
import chess
from oracle import nn_ref
from oracle.encoding_ref import encode_board, get_legal_actions
from matrix0_b200.model import NetConfig, reference_init

BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float32)
G = torch.tensor([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], dtype=torch.float32)
AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float32)
_conv2d = F.conv2d     # nn_ref.F is this module: the patch below replaces the attribute for everybody


def rnd(t, fmt):
    return t.to(fmt).float()


def conv_direct(x, w, fmt):
    return _conv2d(rnd(x, fmt), rnd(w, fmt), padding=1)


def conv_winograd(x, w, fmt):
    B, C = x.shape[:2]
    d = F.pad(rnd(x, fmt), (1, 1, 1, 1)).unfold(2, 4, 2).unfold(3, 4, 2)          # [B, C, 4 tiles y, 4 tiles x, 4, 4]
    V = rnd(torch.einsum("ik,bcyxkl,jl->bcyxij", BT, d, BT), fmt)
    U = rnd(torch.einsum("ik,nckl,jl->ncij", G, w, G), fmt)
    M = torch.einsum("ncij,bcyxij->bnyxij", U, V)
    Y = torch.einsum("ik,bnyxkl,jl->bnyxij", AT, M, AT)                            # [B, N, 4, 4, 2, 2]
    return Y.permute(0, 1, 2, 4, 3, 5).reshape(B, w.shape[0], 8, 8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=2304)
    ap.add_argument("--formats", default="fp16,bf16")
    ap.add_argument("--batch", type=int, default=288)
    args = ap.parse_args()
    g = np.load(os.path.join(ROOT, "tests", "golden", "refinit_golden.npz"))
    d = json.loads(str(g["cfg"]))
    cfg = NetConfig(**{k: v for k, v in d.items() if k in set(NetConfig.__dataclass_fields__)})
    torch.manual_seed(0)
    sd = {k: (v if torch.is_tensor(v) else torch.as_tensor(v)) for k, v in reference_init(cfg).items()}
    n = args.positions
    boards = [chess.Board(str(f)) for f in g["fens"][:n]]
    x = torch.from_numpy(np.stack([encode_board(b) for b in boards]))
    legal = torch.from_numpy(np.stack([get_legal_actions(b) for b in boards]))
    real_conv2d = _conv2d
    modes = [("fp32", "direct")] + [(f, m) for f in args.formats.split(",") for m in ("direct", "winograd")]
    for fmt_name, mode in modes:
        fmt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[fmt_name]

        def patched(inp, weight, *a, **kw):
            if fmt_name != "fp32" and weight.shape[2:] == (3, 3) and weight.shape[0] == weight.shape[1] == cfg.channels and kw.get("padding") == 1:
                return (conv_direct if mode == "direct" else conv_winograd)(inp, weight, fmt)
            return real_conv2d(inp, weight, *a, **kw)

        nn_ref.F.conv2d = patched
        try:
            ps, vs = [], []
            with torch.no_grad():
                for i in range(0, n, args.batch):
                    p, v = nn_ref.forward(sd, cfg, x[i:i + args.batch])
                    ps.append(p)
                    vs.append(v)
        finally:
            nn_ref.F.conv2d = real_conv2d
        p, v = torch.cat(ps), torch.cat(vs)
        top1 = p.argmax(1).numpy()
        top1l = torch.where(legal.bool(), p, torch.full_like(p, -1e30)).argmax(1).numpy()
        print(json.dumps({"operands": fmt_name, "tower_convolutions": mode, "n": n,
                          "top1_vs_reference_fp32": float((top1 == g["top1"][:n]).mean()),
                          "top1_legal_vs_reference_fp32": float((top1l == g["top1_legal"][:n]).mean()),
                          "max_dv_vs_reference_fp32": float(np.abs(v.numpy() - g["value"][:n]).max())}), flush=True)


if __name__ == "__main__":
    main()
