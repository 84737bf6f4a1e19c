"""FourCastNet (AFNONet) on the B200 kernels: drop-ins for the reference's ``Mlp``, ``Block``, ``PatchEmbed`` and
``AFNONet`` (src/dlwpbench/models/fourcastnet/fourcastnet.py:42-57, 156-193, 214-361, 530-543 and the nsbench copy
src/nsbench/models/fourcastnet/fourcastnet.py:40-55, 129-165, 185-300, 303-317) -- same constructors, same
parameters / ``state_dict`` keys, same ``forward``.

Everything between the input frame and the output frame runs in the C-ABI kernels:

    PatchEmbed (patch x patch conv = a Linear over im2col tokens) + bias + pos_embed   tcgen05 GEMM, fused epilogue
    Block:  LN1 -> AFNO2D (+ its own residual + the block's skip, fused into the row synthesis)
            LN2 -> fc1 + bias + GELU -> fc2 + bias + skip                               LayerNorm kernels, tcgen05 GEMMs
    head (Linear, no bias)                                                              tcgen05 GEMM
    backward of all of the above (data gradients with the weights used transposed straight from HBM, weight gradients
    as split-K GEMMs over the token axis, LayerNorm backward with the skip gradient folded in)

torch is used for the im2col / pixel-shuffle index permutations of patch sizes > 1 (pure copies) and for autograd
bookkeeping.  There is no CPU path.
"""
from __future__ import annotations

from functools import partial
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .afno import AFNO2D
from .afno_fn import afno_forward, afno_backward

LN_EPS = 1e-6


def _f32c(t):
    return t.contiguous().float()


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) + resid  on tokens x [T, K]; W [N, K] (nn.Linear layout); ``resid`` [T, N] or, with
    ``res_rows`` > 0, [res_rows, N] broadcast over the batch (pos_embed).  ``act``: GELU on / off."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, W, bias, act: bool, resid, res_rows: int, grad_mode: bool):
        if not x.is_cuda:
            raise _lib.SpectralB200Error("FourCastNet(B200) got a CPU tensor: there is no CPU path")
        x, W = _f32c(x), _f32c(W)
        b = _f32c(bias) if bias is not None else None
        r = _f32c(resid).reshape(-1, W.shape[0]) if resid is not None else None
        need = bool(grad_mode) and any(ctx.needs_input_grad)
        if act:
            y, z = ops.gemm(x, W, bias=b, act=1, resid=r, res_rows=res_rows, want_z=need) if need else \
                (ops.gemm(x, W, bias=b, act=1, resid=r, res_rows=res_rows), None)
        else:
            y, z = ops.gemm(x, W, bias=b, resid=r, res_rows=res_rows), None
        ctx.act, ctx.res_rows = bool(act), int(res_rows)
        ctx.has_bias, ctx.has_res = b is not None, r is not None
        ctx.res_shape = resid.shape if resid is not None else None
        ctx.sinks = (ops.grad_sink(W), ops.grad_sink(bias))
        ctx.save_for_backward(x, W, z)
        return y

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, gy):
        x, W, z = ctx.saved_tensors
        gy = _f32c(gy)
        T, N = gy.shape
        if ctx.act:
            # dz = gy * GELU'(z): an identity-weight GEMM would waste the tensor cores; the GELU' product is fused into the
            # data-gradient GEMM below instead (act = 2 multiplies ITS output, so apply it to gy first with a plain kernel)
            gz = ops.gelu_bwd(gy, z)
        else:
            gz = gy
        gx = gW = gb = gres = None
        if ctx.needs_input_grad[0]:
            gx = ops.gemm(gz, W, b_mn=True)                                   # [T, K] = gz [T, N] . W [N, K]
        if ctx.needs_input_grad[1]:
            gW = ops.gemm(gz, x, a_mn=True, b_mn=True, split_k=True, out=ctx.sinks[0])     # [N, K] = gz^T x
            if ctx.sinks[0] is not None:
                gW = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = ops.colsum(gz, out=ctx.sinks[1])
            if ctx.sinks[1] is not None:
                gb = None
        if ctx.has_res and ctx.needs_input_grad[4]:
            if ctx.res_rows > 0:
                gres = ops.batch_sum(gy.reshape(T // ctx.res_rows, ctx.res_rows * N)).reshape(ctx.res_shape)
            else:
                gres = gy.reshape(ctx.res_shape)
        return gx, gW, gb, None, gres, None, None


class BlockFn(torch.autograd.Function):
    """One FourCastNet block (reference ``Block.forward``, dlwpbench :181-193):

        x1 = AFNO2D(LN1(x)) [+ x]            (AFNO2D adds its own input; ``+ x`` when double_skip)
        y  = fc2(GELU(fc1(LN2(x1)))) + (x1 if double_skip else x)

    args: x [B,h,w,C], g1, be1, w1, b1, w2, b2 (AFNO2D), g2, be2, fc1_w, fc1_b, fc2_w, fc2_b, nb, lam, frac, double_skip,
    grad_mode."""

    @staticmethod
    @_lib.on_tensor_device
    def forward(ctx, x, g1, be1, aw1, ab1, aw2, ab2, g2, be2, fw1, fb1, fw2, fb2, nb, lam, frac, double_skip, grad_mode):
        if not x.is_cuda:
            raise _lib.SpectralB200Error("FourCastNet Block (B200) got a CPU tensor: there is no CPU path")
        x = _f32c(x)
        B, h, w, C = x.shape
        T = B * h * w
        g1c, be1c, g2c, be2c = (_f32c(t) for t in (g1, be1, g2, be2))
        aw1c, ab1c, aw2c, ab2c = (_f32c(t) for t in (aw1, ab1, aw2, ab2))
        fw1c, fb1c, fw2c, fb2c = (_f32c(t) for t in (fw1, fb1, fw2, fb2))
        need = bool(grad_mode) and any(ctx.needs_input_grad)
        x2 = x.view(T, C)
        xn1, m1, r1 = ops.layernorm_fwd(x2, g1c, be1c, LN_EPS)
        xn1_4 = xn1.view(B, h, w, C)
        x1_4, asaved = afno_forward(xn1_4, aw1c, ab1c, aw2c, ab2c, int(nb), float(lam), float(frac), xn1_4,
                                    x if double_skip else None)
        x1 = x1_4.view(T, C)
        xn2, m2, r2 = ops.layernorm_fwd(x1, g2c, be2c, LN_EPS)
        # fc1 writes only the pre-activation z1; h1 = GELU(z1) is formed on chip by the consumers' split pass (fc2 below,
        # the fc2 weight gradient in the backward), so the activated [T, 4C] tensor never exists in HBM
        z1 = ops.gemm(xn2, fw1c, bias=fb1c, z_only=True)
        y = ops.gemm(z1, fw2c, bias=fb2c, resid=x1 if double_skip else x2, a_gelu=True)
        if need:
            ctx.double_skip = bool(double_skip)
            ctx.dims = (B, h, w, C)
            ctx.amisc = asaved[:2]
            ctx.sinks = [ops.grad_sink(t) for t in (g1, be1, g2, be2, fw1, fb1, fw2, fb2)]
            ctx.save_for_backward(x2, m1, r1, x1, m2, r2, xn2, z1, asaved[2], asaved[3], asaved[4],
                                  g1c, g2c, aw1c, aw2c, fw1c, fw2c)
        return y.view(B, h, w, C)

    @staticmethod
    @_lib.on_tensor_device
    def backward(ctx, gy):
        (x2, m1, r1, x1, m2, r2, xn2, z1, Xh, O1, Yh, g1c, g2c, aw1c, aw2c, fw1c, fw2c) = ctx.saved_tensors
        B, h, w, C = ctx.dims
        T = B * h * w
        s_g1, s_be1, s_g2, s_be2, s_fw1, s_fb1, s_fw2, s_fb2 = ctx.sinks
        gy2 = _f32c(gy).view(T, C)
        keep = lambda t, s: None if s is not None else t
        # ---- token MLP ----
        gfw2 = ops.gemm(gy2, z1, a_mn=True, b_mn=True, b_gelu=True, split_k=True, out=s_fw2)   # [C, 4C] = gy^T GELU(z1)
        gfb2 = ops.colsum(gy2, out=s_fb2)
        gz1 = ops.gemm(gy2, fw2c, b_mn=True, act=2, aux=z1)                                # [T, 4C] = (gy W2) * GELU'(z1)
        gfw1 = ops.gemm(gz1, xn2, a_mn=True, b_mn=True, split_k=True, out=s_fw1)           # [4C, C]
        gfb1 = ops.colsum(gz1, out=s_fb1)
        gxn2 = ops.gemm(gz1, fw1c, b_mn=True)                                              # [T, C]
        # ---- LN2 (+ the skip around the MLP when it starts at x1) ----
        gx1, gg2, gbe2 = ops.layernorm_bwd(gxn2, x1, g2c, m2, r2, dres=gy2 if ctx.double_skip else None,
                                           out_g=s_g2, out_b=s_be2)
        # ---- AFNO2D: x1 = filter(xn1) + xn1 (+ x) ----
        gx1_4 = gx1.view(B, h, w, C)
        gxn1_4, gaw1, gab1, gaw2, gab2 = afno_backward((*ctx.amisc, Xh, O1, Yh), aw1c, aw2c, gx1_4, gx1_4, need_gx=True)
        # ---- LN1 (+ the skip gradient: gx1 with double_skip, else the block-level skip gy) ----
        gx, gg1, gbe1 = ops.layernorm_bwd(gxn1_4.view(T, C), x2, g1c, m1, r1, dres=gx1 if ctx.double_skip else gy2,
                                          out_g=s_g1, out_b=s_be1)
        return (gx.view(B, h, w, C), keep(gg1, s_g1), keep(gbe1, s_be1), gaw1, gab1, gaw2, gab2, keep(gg2, s_g2),
                keep(gbe2, s_be2), keep(gfw1, s_fw1), keep(gfb1, s_fb1), keep(gfw2, s_fw2), keep(gfb2, s_fb2),
                None, None, None, None, None)


# --------------------------------------------------------------------------------------
# modules (parameter containers with the reference's names; forward through the Functions above)
# --------------------------------------------------------------------------------------
class Mlp(nn.Module):
    """reference ``Mlp`` (:42-57): Linear -> GELU -> Linear (dropout 0 only)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if act_layer is not nn.GELU:
            raise NotImplementedError("Mlp(B200): only nn.GELU is implemented (fused epilogue)")
        if drop:
            raise NotImplementedError("Mlp(B200): dropout is not implemented (every shipped config uses 0)")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        shp = x.shape
        gm = torch.is_grad_enabled()
        h = LinearFn.apply(x.reshape(-1, shp[-1]), self.fc1.weight, self.fc1.bias, True, None, 0, gm)
        y = LinearFn.apply(h, self.fc2.weight, self.fc2.bias, False, None, 0, gm)
        return y.reshape(*shp[:-1], y.shape[-1])


class Block(nn.Module):
    """reference ``Block`` (dlwpbench :156-193 / nsbench :129-165)."""

    def __init__(self, dim, filter=AFNO2D, mlp_ratio=4., drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 double_skip=True, num_blocks=8, sparsity_threshold=0.01, hard_thresholding_fraction=1.0):
        super().__init__()
        if drop_path:
            raise NotImplementedError("Block(B200): drop_path is not implemented (every shipped config uses 0)")
        self.norm1 = norm_layer(dim)
        self.filter = filter(dim, num_blocks, sparsity_threshold, hard_thresholding_fraction)
        if not isinstance(self.filter, AFNO2D):
            raise NotImplementedError("Block(B200): only the AFNO2D filter is implemented")
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.double_skip = double_skip
        for n in (self.norm1, self.norm2):
            if abs(n.eps - LN_EPS) > 1e-12:
                raise NotImplementedError(f"Block(B200): LayerNorm eps must be {LN_EPS} (what AFNONet passes), got {n.eps}")

    def forward(self, x):
        f, m = self.filter, self.mlp
        dtype = x.dtype
        y = BlockFn.apply(x, self.norm1.weight, self.norm1.bias, f.w1, f.b1, f.w2, f.b2, self.norm2.weight, self.norm2.bias,
                          m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, f.num_blocks, float(f.sparsity_threshold),
                          float(f.hard_thresholding_fraction), bool(self.double_skip), torch.is_grad_enabled())
        return y.type(dtype)


class PatchEmbed(nn.Module):
    """reference ``PatchEmbed`` (:530-543): Conv2d(kernel = stride = patch) -> [B, num_patches, embed_dim]."""

    def __init__(self, img_size=(224, 224), patch_size=(16, 16), in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = tuple(img_size), tuple(patch_size)
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0])
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def tokens(self, x):
        """im2col: [B,C,H,W] -> [B*h*w, C*p1*p2] (K padded to a multiple of 4 floats for the TMA row stride)."""
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        p1, p2 = self.patch_size
        h, w = H // p1, W // p2
        t = x.float().reshape(B, C, h, p1, w, p2).permute(0, 2, 4, 1, 3, 5).reshape(B * h * w, C * p1 * p2)
        pad = (-t.shape[1]) % 4
        return (F.pad(t, (0, pad)) if pad else t).contiguous(), pad

    def forward(self, x, pos_embed=None):
        t, pad = self.tokens(x)
        Wm = self.proj.weight.reshape(self.proj.out_channels, -1)
        if pad:
            Wm = F.pad(Wm, (0, pad))
        y = LinearFn.apply(t, Wm, self.proj.bias, False, pos_embed, self.num_patches if pos_embed is not None else 0,
                           torch.is_grad_enabled())
        return y.reshape(x.shape[0], self.num_patches, -1)


class _AFNOCore(nn.Module):
    """Shared body of the two AFNONet flavours: patch_embed, pos_embed, blocks, (unused) norm, head."""

    def _build(self, in_chans, embed_dim, depth, mlp_ratio, drop_rate, drop_path_rate, num_blocks, sparsity_threshold,
               hard_thresholding_fraction, use_pos_embed):
        if drop_rate or drop_path_rate:
            raise NotImplementedError("AFNONet(B200): dropout / drop_path are not implemented (shipped configs use 0)")
        norm_layer = partial(nn.LayerNorm, eps=LN_EPS)
        self.patch_embed = PatchEmbed(img_size=self.img_size, patch_size=self.patch_size, in_chans=in_chans, embed_dim=embed_dim)
        num_patches = self.patch_embed.num_patches
        self.use_pos_embed = use_pos_embed
        if use_pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.h = self.img_size[0] // self.patch_size[0]
        self.w = self.img_size[1] // self.patch_size[1]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, mlp_ratio=mlp_ratio, drop=drop_rate, drop_path=0., norm_layer=norm_layer,
                  num_blocks=self.num_blocks, sparsity_threshold=sparsity_threshold,
                  hard_thresholding_fraction=hard_thresholding_fraction) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)          # present in the reference's state_dict, never used in its forward
        self.head = nn.Linear(embed_dim, self.out_chans * self.patch_size[0] * self.patch_size[1], bias=False)
        if use_pos_embed:
            nn.init.trunc_normal_(self.pos_embed, std=.02)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward_features(self, x):
        B = x.shape[0]
        x = self.patch_embed(x, self.pos_embed if self.use_pos_embed else None)
        x = x.reshape(B, self.h, self.w, self.embed_dim)
        for blk in self.blocks:
            x = blk(x)
        return x

    def step(self, x_t):
        """[B, C_in, H, W] -> [B, C_out, H, W]: forward_features, head, pixel shuffle (reference :343-353)."""
        B = x_t.shape[0]
        f = self.forward_features(x_t)
        y = LinearFn.apply(f.reshape(-1, self.embed_dim), self.head.weight, None, False, None, 0, torch.is_grad_enabled())
        p1, p2 = self.patch_size
        y = y.reshape(B, self.h, self.w, p1, p2, self.out_chans).permute(0, 5, 1, 3, 2, 4)
        return y.reshape(B, self.out_chans, self.h * p1, self.w * p2)


class AFNONet(_AFNOCore):
    """dlwpbench flavour (src/dlwpbench/models/fourcastnet/fourcastnet.py:214-361): inputs are the constant /
    prescribed / prognostic tensors; ``forward`` runs the autoregressive loop of ``dlwp_sequence_forward``."""

    def __init__(self, img_height=720, img_width=1440, patch_size=(16, 16), constant_channels: int = 4,
                 prescribed_channels: int = 0, prognostic_channels: int = 1, filter="AFNO2D", embed_dim=768, depth=12,
                 mlp_ratio=4., drop_rate=0., drop_path_rate=0., num_blocks=16, sparsity_threshold=0.01,
                 hard_thresholding_fraction=1.0, context_size: int = 1, use_pos_embed: bool = True, **kwargs):
        super().__init__()
        if filter != "AFNO2D":
            raise NotImplementedError("AFNONet(B200): only filter='AFNO2D' is implemented")
        self.img_size = (img_height, img_width)
        self.patch_size = tuple(patch_size)
        self.in_chans = constant_channels + (prescribed_channels + prognostic_channels) * context_size
        self.out_chans = prognostic_channels
        self.num_features = self.embed_dim = embed_dim
        self.num_blocks = num_blocks
        self.context_size = context_size
        self._build(self.in_chans, embed_dim, depth, mlp_ratio, drop_rate, drop_path_rate, num_blocks, sparsity_threshold,
                    hard_thresholding_fraction, use_pos_embed)

    def forward(self, constants=None, prescribed=None, prognostic=None):
        from .rollout import dlwp_sequence_forward
        # out_t = prognostic_t[:, -1] + step(x_t): the loop adds the residual, ``step`` is the network proper
        return dlwp_sequence_forward(self.step, constants, prescribed, prognostic, self.context_size)


class AFNONetNS(_AFNOCore):
    """nsbench flavour (src/nsbench/models/fourcastnet/fourcastnet.py:185-300): ``forward(x [B,T,D,H,W],
    teacher_forcing_steps)`` with the reference's teacher-forcing / closed-loop window logic."""

    def __init__(self, img_height=720, img_width=1440, patch_size=(16, 16), in_chans=2, out_chans=2, embed_dim=768, depth=12,
                 mlp_ratio=4., drop_rate=0., drop_path_rate=0., num_blocks=16, sparsity_threshold=0.01,
                 hard_thresholding_fraction=1.0, context_size: int = 1, **kwargs):
        super().__init__()
        self.img_size = (img_height, img_width)
        self.patch_size = tuple(patch_size)
        self.in_chans = in_chans * context_size
        self.out_chans = out_chans
        self.num_features = self.embed_dim = embed_dim
        self.num_blocks = num_blocks
        self.context_size = context_size
        self._build(self.in_chans, embed_dim, depth, mlp_ratio, drop_rate, drop_path_rate, num_blocks, sparsity_threshold,
                    hard_thresholding_fraction, True)

    def forward(self, x, teacher_forcing_steps: int = 50):
        outs = []
        out = None
        for t in range(x.shape[1]):
            if t < teacher_forcing_steps:
                x_t_in = x[:, max(0, t - (self.context_size - 1)):t + 1]
            elif self.context_size == 0:
                x_t_in = out
            else:
                ts = max(0, (teacher_forcing_steps - t - 1) + self.context_size)
                x_obs = x[:, teacher_forcing_steps - ts:teacher_forcing_steps]
                x_out = torch.stack(outs[-(self.context_size - ts):], dim=1)
                x_t_in = torch.cat([x_obs, x_out], dim=1)
            if t < self.context_size - 1:
                out = x_t_in[:, -1]
            else:
                b, tt, d, hh, ww = x_t_in.shape
                out = x_t_in[:, -1] + self.step(x_t_in.reshape(b, tt * d, hh, ww))
            outs.append(out)
        return torch.stack(outs, dim=1)
