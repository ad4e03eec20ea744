"""TEST INFRASTRUCTURE: a torch (fp64, CPU) statement of the operations ``unimm_b200.train_ops.DeviceOps`` offers, with the same
in-place / accumulate / "written to the gradient view" semantics.  ``tests/test_train_step_cpu.py`` runs
``unimm_b200.train_step.TrainStep`` over it to check the ORCHESTRATION of the training step (layer schedule in reverse, which saved
tensor feeds which backward, fused Q|K|V views of the flat parameter buffer, padded tensors, parameter groups) against
``torch.autograd`` of the oracle — without a GPU.  The product never imports this file; the kernels themselves are checked against
autograd on the GPU (tests/test_train_step_gpu.py)."""
from __future__ import annotations

import math

import numpy as np
import torch

from unimm_b200.descriptors import dense_co_mask, dense_text_mask
from unimm_b200.train_ops import ACT_GELU, ACT_RELU, EW_ADD, EW_AXPY, EW_MUL, EW_RELU_BWD, EW_SCALE, MASK_CO_INTERVAL, MASK_KEY_VECTOR, MASK_TEXT_SELF

DT = torch.float64


def _lowbias32(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def keep_mask(drop, shape):
    """The keep-mask the device kernels regenerate (csrc/common.cuh: drop_keep): element i of the flattened tensor is kept iff
    lowbias32(lowbias32(i ^ seed) + seed) >= p * 2^32; kept values are scaled by 1 / (1 - p)."""
    seed, p = drop
    n = int(np.prod(shape))
    assert n < 2 ** 32
    idx = np.arange(n, dtype=np.uint64)
    h = _lowbias32((_lowbias32(idx ^ np.uint64(seed)) + np.uint64(seed)) & np.uint64(0xFFFFFFFF))
    keep = h >= np.uint64(min(int(p * 4294967296.0), 0xFFFFFFFF))
    return torch.from_numpy(keep.reshape(shape)).to(DT) / (1.0 - p)


def _gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


class TorchOps:
    precision = "fp64"

    def join_side(self):            # DeviceOps orders its second (wgrad) stream here; one stream in this restatement
        pass

    def begin_step(self):
        pass

    def new_amax_cell(self):
        return None

    def register_amax(self, t, cell):
        pass

    def empty32(self, *shape):
        return torch.full(shape, float("nan"), dtype=DT)       # reading an "empty" buffer before it is written poisons the result

    def zeros32(self, *shape):
        return torch.zeros(*shape, dtype=DT)

    def empty16(self, *shape):
        return torch.full(shape, float("nan"), dtype=DT)

    def to_lp(self, x):
        return x.clone()

    def cast_into(self, x, out16):
        out16.copy_(x)

    def ew(self, op, a, b=None, out=None, alpha=1.0):
        out = a if out is None else out
        r = {EW_ADD: lambda: a + b, EW_MUL: lambda: a * b, EW_RELU_BWD: lambda: a * (b > 0), EW_SCALE: lambda: alpha * a,
             EW_AXPY: lambda: a + alpha * b}[op]()
        out.copy_(r)
        return out

    def mul(self, a, b):
        return a * b

    def relu_backward(self, dy, y):
        dy.mul_((y > 0).to(DT))
        return dy

    def gather_rows(self, src, idx):
        return src[idx.long()].clone()

    def scatter_add_rows(self, src, idx, dst):
        dst.index_add_(0, idx.long(), src)

    def embed_text_sum(self, ids, seg, pos, word, pos_emb, type_emb, type_ext, type_vocab):
        ids, seg, pos = ids.reshape(-1), seg.reshape(-1), pos.reshape(-1)
        is_ext = seg >= type_vocab
        ty = torch.where(is_ext.unsqueeze(-1), type_ext[(seg - type_vocab).clamp(min=0)], type_emb[seg.clamp(max=type_vocab - 1)])
        return word[ids] + pos_emb[pos] + ty

    def embed_text_backward(self, dsum, ids, seg, pos, g_word, g_pos, g_type, g_type_ext, type_vocab):
        ids, seg, pos = ids.reshape(-1), seg.reshape(-1), pos.reshape(-1)
        g_word.index_add_(0, ids, dsum)
        g_pos.index_add_(0, pos, dsum)
        is_ext = seg >= type_vocab
        g_type.index_add_(0, seg[~is_ext], dsum[~is_ext])
        g_type_ext.index_add_(0, seg[is_ext] - type_vocab, dsum[is_ext])

    def layernorm(self, x, gamma, beta, want32=True, want16=True):
        y = torch.nn.functional.layer_norm(x, (x.shape[-1],), gamma, beta, 1e-12)
        return (y if want32 else None), (y.clone() if want16 else None)

    def layernorm_backward(self, dy, x, gamma, g_gamma, g_beta):
        xx = x.detach().clone().requires_grad_()
        gg, bb = gamma.detach().clone().requires_grad_(), torch.zeros_like(gamma).requires_grad_()
        y = torch.nn.functional.layer_norm(xx, (x.shape[-1],), gg, bb, 1e-12)
        dx, dg, db = torch.autograd.grad(y, (xx, gg, bb), dy)
        g_gamma.copy_(dg)
        g_beta.copy_(db)
        return dx

    def gelu(self, t, want32=False, want16=True):
        g = _gelu(t)
        return (g if want32 else None), (g.clone() if want16 else None)

    def gelu_backward(self, dy, t):
        tt = t.detach().clone().requires_grad_()
        (d,) = torch.autograd.grad(_gelu(tt), tt, dy.clone())
        dy.copy_(d)
        return dy

    def dropout(self, x, drop, want16=True):
        y = x * keep_mask(drop, x.shape)
        return y, (y.clone() if want16 else None)

    def dropout_backward(self, dy, drop):
        dy.mul_(keep_mask(drop, dy.shape))
        return dy

    def linear(self, x, w, bias, residual=None, act=0, want32=True, want16=False, pre_act32=False, drop=None):
        y = x @ w.t() + bias
        if drop is not None:
            y = y * keep_mask(drop, y.shape)
        if pre_act32:
            return y, (_gelu(y) if act == ACT_GELU else torch.relu(y))
        if act == ACT_GELU:
            y = _gelu(y)
        elif act == ACT_RELU:
            y = torch.relu(y)
        if residual is not None:
            y = y + residual
        return (y if want32 else None), (y.clone() if want16 else None)

    def linear_f32(self, x, w, bias, residual=None, act=0):
        return self.linear(x, w, bias, residual, act)[0]

    def linear_backward(self, dy, x, w, g_w, g_b, need_dx=True, dx_accum=None, gelu_t=None, dx_amax=False, drop=None):
        assert tuple(g_w.shape) == (dy.shape[1], x.shape[1])
        if drop is not None:
            dy = dy * keep_mask(drop, dy.shape)
        if gelu_t is not None:
            dy = self.gelu_backward(dy.clone(), gelu_t)
        g_w.copy_(dy.t() @ x)
        if g_b is not None:
            g_b.copy_(dy.sum(0))
        if not need_dx:
            return None
        dx = dy @ w
        if dx_accum is not None:
            dx_accum.add_(dx)
            return dx_accum
        return dx

    @staticmethod
    def _add_mask(B, Sq, Skv, mask_kind, desc, key_mask):
        if mask_kind == MASK_TEXT_SELF:
            m = dense_text_mask(desc, Skv).to(DT)[:, None]                        # [B,1,S,S]
        elif mask_kind == MASK_CO_INTERVAL:
            m = dense_co_mask(desc, Skv).to(DT)[:, None, None, :].expand(B, 1, Sq, Skv)
        else:
            m = key_mask.to(DT)[:, None, None, :].expand(B, 1, Sq, Skv)
        return (1.0 - m) * -10000.0

    def _attn(self, q, k, v, B, heads, D, Sq, Skv, add, drop=None):
        qh = q.reshape(B, Sq, heads, D).permute(0, 2, 1, 3)
        kh = k.reshape(B, Skv, heads, D).permute(0, 2, 1, 3)
        vh = v.reshape(B, Skv, heads, D).permute(0, 2, 1, 3)
        s = qh @ kh.transpose(-1, -2) / math.sqrt(D) + add
        pr = torch.softmax(s, -1)
        if drop is not None:
            pr = pr * keep_mask(drop, (B, heads, Sq, Skv))
        o = pr @ vh
        return o.permute(0, 2, 1, 3).reshape(B * Sq, heads * D), torch.logsumexp(s, -1)

    def attention(self, q, k, v, B, heads, D, Sq, Skv, mask_kind, desc=None, key_mask=None, drop=None):
        return self._attn(q, k, v, B, heads, D, Sq, Skv, self._add_mask(B, Sq, Skv, mask_kind, desc, key_mask), drop)

    def attention_backward(self, q, k, v, o, lse, dO, B, heads, D, Sq, Skv, mask_kind, desc, key_mask, dq, dk, dv, amax_cell=None, drop=None):
        qq, kk, vv = (t.detach().clone().requires_grad_() for t in (q, k, v))
        oo, _ = self._attn(qq, kk, vv, B, heads, D, Sq, Skv, self._add_mask(B, Sq, Skv, mask_kind, desc, key_mask), drop)
        a, b, c = torch.autograd.grad(oo, (qq, kk, vv), dO)
        dq.copy_(a)
        dk.copy_(b)
        dv.copy_(c)

    def lm_head_loss_backward(self, h, e, bias, labels, weight, grad_scale, g_e, g_bias):
        hh, ee, bb = h.detach().clone().requires_grad_(), e.detach().clone().requires_grad_(), bias.detach().clone().requires_grad_()
        logits = hh @ ee.t() + bb
        logp = torch.log_softmax(logits, -1).gather(1, labels.long()[:, None])[:, 0]
        ul = torch.log(torch.clamp(1.0 - torch.softmax(logits, -1), min=1e-6)).gather(1, labels.long()[:, None])[:, 0]
        loss = (-(logp * weight)[weight > 0]).sum() + (-ul[weight == -1]).sum()
        dH, dE, db = torch.autograd.grad(loss * grad_scale, (hh, ee, bb))
        g_e.copy_(dE)
        g_bias.copy_(db)
        return dH, logp.detach()

    def lm_ul_value(self, logp, weight, scale):
        ul = torch.log(torch.clamp(1.0 - torch.exp(logp), min=1e-6))
        return ((-(logp * weight)[weight > 0]).sum() + (-ul[weight == -1]).sum()).reshape(1) * scale

    def nsp_ce(self, logits, labels, nsp_weight, grad_scale):
        x = logits.detach().clone().requires_grad_()
        w = torch.ones(2, dtype=DT) if nsp_weight is None else (nsp_weight / nsp_weight[0]).to(DT)
        loss = torch.nn.functional.cross_entropy(x, labels, weight=w, reduction="mean")
        (d,) = torch.autograd.grad(loss * grad_scale, x)
        return loss.detach().reshape(1), d

    def image_kl(self, logits, C_real, target, target_row, image_label, grad_scale):
        x = logits.detach().clone().requires_grad_()
        t = target[target_row.long()]
        kl = torch.nn.functional.kl_div(torch.log_softmax(x[:, :C_real], -1), t, reduction="none")
        sel = (image_label == 1)
        loss = (kl * sel[:, None].to(DT)).sum() / sel.sum()
        (d,) = torch.autograd.grad(loss * grad_scale, x)
        return loss.detach().reshape(1), d

    def nsp_prob0(self, logits):
        return torch.softmax(logits, -1)[:, 0].clone()

    def nsp_prob0_backward(self, logits, dp0, dlogits_accum):
        x = logits.detach().clone().requires_grad_()
        (d,) = torch.autograd.grad(torch.softmax(x, -1)[:, 0], x, dp0)
        dlogits_accum.add_(d)

    def neural_ndcg_backward(self, y_pred, y_true, grad_scale, temperature=1.0, max_iter=50, tol=1e-6):
        from oracle import rank_loss as orl
        g = orl.neural_ndcg_transposed_grad(y_pred.numpy(), y_true.numpy(), temperature, max_iter, tol)
        _, ndcg, _ = orl.neural_ndcg_transposed(y_pred.numpy(), y_true.numpy(), temperature, max_iter, tol)
        return torch.from_numpy(g).to(DT) * grad_scale, torch.from_numpy(ndcg).to(DT)

    def adamw(self, p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, correct_bias, inv_grad_scale, p16):
        gi = g * inv_grad_scale
        m.mul_(beta1).add_(gi, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
        step_size = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step) if correct_bias else lr
        p.addcdiv_(m, v.sqrt() + eps, value=-step_size)
        if weight_decay > 0:
            p.add_(p, alpha=-lr * weight_decay)
        if p16 is not None:
            p16.copy_(p)
