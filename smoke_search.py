"""Helper module of __graft_entry__.smoke() (not part of the product package: it checks against the oracle).  smoke(): one tiny closed-loop invocation of the hot path on cuda:0 checked against the oracle:
a small random-weight evaluator, four games, one 96-simulation search, visit counts == oracle."""
from __future__ import annotations


def run() -> None:
    import numpy as np
    import torch
    from oracle import chess_shim  # noqa: F401
    import chess
    from oracle import nn_ref
    from oracle.mcts_ref import RefConfig, RefMCTS
    from matrix0_b200.model import NetConfig, PolicyValueNet, parameter_shapes
    from matrix0_b200.selfplay import SelfPlayEngine
    cfg = NetConfig(channels=32, blocks=3, attention_heads=2, policy_factor_rank=16, norm="group", activation="silu",
                    value_activation="leaky_relu", preact=True, infer_attention_stride=1, ssl_tasks=["piece"])
    sd = nn_ref.make_state_dict(parameter_shapes(cfg), seed=3)
    net = PolicyValueNet(cfg, device="cuda:0", precision="fp32")
    net.load_state_dict(sd, strict=True)
    # evaluator vs the fp32 oracle
    x = torch.rand(3, 19, 8, 8, generator=torch.Generator().manual_seed(0))
    p, v = net.forward(x)
    with torch.no_grad():
        pr, vr = nn_ref.forward(sd, cfg, x)
    assert float((p.cpu() - pr).abs().max()) <= 1e-4 * max(1.0, float(pr.abs().max())), "evaluator mismatch"
    assert float((v.cpu() - vr).abs().max()) <= 1e-4
    tensor_core_leg()
    # search vs the oracle search with the same evaluator
    kw = dict(cpuct=2.5, cpuct_start=3.0, cpuct_end=2.0, cpuct_plies=40, fpu_reduction=0.1, draw_penalty=-0.05, legal_softmax=True,
              selection_jitter=0.05, inference_batch_size=32, num_simulations=96)
    boards = [chess.Board(), chess.Board("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"),
              chess.Board("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1"), chess.Board("6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1")]
    sp = SelfPlayEngine(net, {"mcts": kw, "selfplay": {"num_simulations": 96}}, games=len(boards), device=0, deterministic=True, precision="fp32")
    sp.engine.set_boards(boards)
    sp.begin_move()
    for _ in range(sp.batches_per_move()):
        sp.search_step()
    sp.engine.result(with_pi=False)
    cnt = sp.engine.res_count.cpu().numpy()
    mv = sp.engine.res_moves.cpu().numpy().view(np.uint16)
    vis = sp.engine.res_visits.cpu().numpy()
    for g, b in enumerate(boards):
        ref = RefMCTS(RefConfig(dirichlet_frac=0.0, enable_entropy_noise=False, **kw), net, jitter_value=0.5)
        vc, _, _ = ref.run(b.copy(), ply=0)
        exp = [(m.from_square | (m.to_square << 6) | ((m.promotion or 0) << 12), n) for m, n in vc.items()]
        got = [(int(mv[g, j]), int(vis[g, j])) for j in range(int(cnt[g]))]
        assert got == exp, f"search mismatch for {b.fen()}"


def tensor_core_leg() -> None:
    """The throughput path on cuda:0: one tcgen05 CTA-pair 3x3 convolution (m0_tc_conv) against plain PyTorch fp32 on the same
    16-bit-rounded operands, and one fp16 tensor-core forward of a 64-channel evaluator against the fp32 oracle."""
    import torch
    from oracle import nn_ref
    from matrix0_b200 import _native
    from matrix0_b200.model import NetConfig, PolicyValueNet, parameter_shapes
    lib = _native.lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    boards, cin, n = 4, 64, 64
    act = torch.randn((boards, 8, 8, cin), device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn((n, cin, 3, 3), device="cuda", generator=g) / (9 * cin) ** 0.5).to(torch.bfloat16)
    w = wt.permute(0, 2, 3, 1).reshape(n, 9 * cin).contiguous()
    out = torch.empty((boards * 64, n), dtype=torch.float32, device="cuda")
    _native.check(lib.m0_tc_conv(act.data_ptr(), w.data_ptr(), boards, cin, n, 9, out.data_ptr(), _native.current_stream()), "m0_tc_conv")
    ref = torch.nn.functional.conv2d(act.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1).reshape(boards * 64, n)
    assert float((out - ref).abs().max()) <= 2e-3 * max(1.0, float(ref.abs().max())), "tcgen05 convolution mismatch"
    cfg = NetConfig(channels=64, blocks=3, attention_heads=4, policy_factor_rank=32, norm="group", activation="silu",
                    value_activation="leaky_relu", preact=True, infer_attention_stride=1, ssl_tasks=["piece"])
    sd = nn_ref.make_state_dict(parameter_shapes(cfg), seed=4)
    net = PolicyValueNet(cfg, device="cuda:0", precision="fp16")
    net.load_state_dict(sd, strict=True)
    x = (torch.rand(8, 19, 8, 8, generator=torch.Generator().manual_seed(1)) > 0.8).float()
    p, v = net.forward(x)
    with torch.no_grad():
        pr, vr = nn_ref.forward(sd, cfg, x)
    scale = max(1.0, float(pr.abs().max()))
    assert float((p.float().cpu() - pr).abs().max()) <= 3e-2 * scale, "fp16 tensor-core forward mismatch"
    assert float((v.float().cpu() - vr).abs().max()) <= 2e-2, "fp16 tensor-core value mismatch"
