#!/usr/bin/env python
"""Reproducer for the parity note in DESIGN.md section 2: for the NON-HERMITIAN half spectra the SpectralConv
layer produces (the kx = 0 column of the retained block is an arbitrary complex vector, not conjugate-symmetric
in ky), `torch.fft.irfftn` is only defined up to what the backend does with the redundant imaginary parts.

Compared here, on identical inputs, at 64^2 / 128^2 / 256^2 with the reference's retained block:
  cpu        torch.fft.irfftn on the CPU (pocketfft: complex iFFT over H, then C2R over W)   -- the oracle's meaning
  cuda       torch.fft.irfftn on the GPU (cuFFT multi-dimensional C2R)
  cuda_2step ifft over H, then irfft over W, both on the GPU (cuFFT 1-D transforms)
  b200       this repo's kernels (coldft_inv + rowidft), when the library is built
Writes one JSON line per size; run on the GPU box:  python scripts/cufft_vs_cpu_irfftn.py > profiles/r02_cufft_vs_cpu.jsonl
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (torch.linalg.norm(a - b) / torch.linalg.norm(b)).item()


def main():
    torch.manual_seed(0)
    for (H, W, n) in ((64, 64, 16), (128, 128, 32), (256, 256, 32)):
        My, Mx = n, n // 2 + 1
        lo = (H - My) // 2
        B, C = 2, 4
        blk = torch.randn(B, C, My, Mx, dtype=torch.complex64)
        full = torch.zeros(B, C, H, W // 2 + 1, dtype=torch.complex64)
        full[:, :, lo:lo + My, :Mx] = blk                       # shifted layout ...
        full = torch.fft.fftshift(full, dim=-2)                 # ... back to FFT order (neuralop does the same)
        ref64 = torch.fft.irfftn(full.to(torch.complex128), s=(H, W), dim=(-2, -1), norm="forward")
        cpu = torch.fft.irfftn(full, s=(H, W), dim=(-2, -1), norm="forward")
        line = {"H": H, "W": W, "modes": n, "cpu_fp32_vs_cpu_fp64": rel(cpu, ref64)}
        herm = full.clone()
        herm[..., 0] = 0                                         # Hermitian-safe variant: no kx = 0 column at all
        ref_h = torch.fft.irfftn(herm.to(torch.complex128), s=(H, W), dim=(-2, -1), norm="forward")
        if torch.cuda.is_available():
            g = full.cuda()
            cuda = torch.fft.irfftn(g, s=(H, W), dim=(-2, -1), norm="forward")
            two = torch.fft.irfft(torch.fft.ifft(g, dim=-2, norm="forward"), n=W, dim=-1, norm="forward")
            line["cuda_irfftn_vs_cpu_fp64"] = rel(cuda, ref64)
            line["cuda_2step_vs_cpu_fp64"] = rel(two, ref64)
            line["cuda_irfftn_vs_cpu_fp64_without_kx0_column"] = rel(
                torch.fft.irfftn(herm.cuda(), s=(H, W), dim=(-2, -1), norm="forward"), ref_h)
            try:
                from dlwp_benchmark_b200 import ops
                from dlwp_benchmark_b200.plan import fno_plan
                plan = fno_plan(torch.device("cuda"), H, W, (My, Mx))
                Yh = torch.view_as_real(blk).contiguous().cuda()
                Phi = ops.coldft_inv(plan, 0, Yh)
                y, _ = ops.rowidft_pointwise(plan, 0, Phi, None, None, C, 1, None, None, B, C, C, 0, False)
                line["b200_kernels_vs_cpu_fp64"] = rel(y, ref64)
            except Exception as ex:  # noqa: BLE001
                line["b200_kernels_error"] = repr(ex)[:200]
        # the oracle's SpectralConv forward + backward (autograd through rfftn / irfftn) on CUDA vs on the CPU
        try:
            from oracle import spectral_oracle as so
            x = torch.randn(2, 4, H, W)
            w = torch.randn(4, 4, My, Mx, dtype=torch.complex64) * 0.3
            res = {}
            for dev in (["cpu", "cuda"] if torch.cuda.is_available() else ["cpu"]):
                xd, wd = x.detach().to(dev).requires_grad_(True), w.detach().clone().to(dev).requires_grad_(True)
                y = so.spectral_conv_dense(xd, wd, None, [My, Mx])
                y.square().sum().backward()
                res[dev] = (y.detach(), xd.grad, torch.view_as_real(wd.grad))
            if "cuda" in res:
                line["oracle_layer_cuda_vs_cpu(y,gx,gW)"] = [rel(a, b) for a, b in zip(res["cuda"], res["cpu"])]
        except Exception as ex:  # noqa: BLE001
            line["oracle_layer_error"] = repr(ex)[:200]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
