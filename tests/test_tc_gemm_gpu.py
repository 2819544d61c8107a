"""The tcgen05 / TMA implicit-GEMM kernel on its own, against plain PyTorch fp32 references computed from the
same bf16-rounded operands (so the only difference is accumulation order: tolerance 2e-3 relative to the output scale)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_tc(act, w, taps):
    from matrix0_b200 import _native
    lib = _native.lib()
    boards, cin, n = act.shape[0], act.shape[-1], w.shape[0]
    out = torch.empty((boards * 64, n), dtype=torch.float32, device="cuda")
    _native.check(lib.m0_tc_conv(act.data_ptr(), w.data_ptr(), boards, cin, n, taps, out.data_ptr(), _native.current_stream()), "m0_tc_conv")
    torch.cuda.synchronize()
    return out


# n % 32 == 0 runs on the CTA-pair kernel (tc_conv_pair.cuh), other widths on the single-CTA kernel (tc_gemm.cuh)
@pytest.mark.parametrize("boards,cin,n", [(2, 64, 64), (2, 320, 320), (6, 128, 160), (298, 320, 320), (2, 320, 64), (4, 64, 256),
                                           (4, 128, 48), (2, 320, 80), (298, 320, 176), (64, 3840, 80)])
def test_plain_gemm(boards, cin, n):
    g = torch.Generator(device="cuda").manual_seed(boards * 1000 + cin + n)
    act = torch.randn((boards, 64, cin), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((n, cin), device="cuda", generator=g) / cin ** 0.5).to(torch.bfloat16)
    out = run_tc(act, w, 1)
    ref = act.float().reshape(-1, cin) @ w.float().t()
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("boards,cin,n", [(2, 64, 64), (2, 320, 320), (4, 128, 320), (300, 320, 320)])
def test_conv3x3(boards, cin, n):
    g = torch.Generator(device="cuda").manual_seed(7 + boards + cin + n)
    act = torch.randn((boards, 8, 8, cin), device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn((n, cin, 3, 3), device="cuda", generator=g) / (9 * cin) ** 0.5).to(torch.bfloat16)
    w = wt.permute(0, 2, 3, 1).reshape(n, 9 * cin).contiguous()          # k = (ky*3+kx)*cin + ci
    out = run_tc(act, w, 9)
    ref = torch.nn.functional.conv2d(act.float().permute(0, 3, 1, 2), wt.float(), padding=1)   # plain PyTorch fp32 reference
    ref = ref.permute(0, 2, 3, 1).reshape(boards * 64, n)
    err = (out - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err
