"""Device operations of the training step (``unimm_b200/train_step.py``): thin, typed wrappers over the kernel-level C ABI
(``unimm_k_*`` / ``unimm_t_*`` of ``include/unimm_b200.h``).  PyTorch is used for device memory and the stream only; every
arithmetic operation below is one of this library's CUDA kernels.  There is no CPU path: the constructor refuses a non-CUDA device.

Conventions: ``x32`` = fp32 tensor, ``x16`` = 16-bit GEMM operand (fp16 or bf16 according to ``precision``); matrices are 2-D,
row-major, possibly column slices of a wider matrix (the leading dimension is taken from ``stride(0)``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from ._lib import LP_BF16, LP_FP16, check, lib, ptr

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
MASK_TEXT_SELF, MASK_KEY_VECTOR, MASK_CO_INTERVAL = 0, 1, 2
EW_ADD, EW_MUL, EW_RELU_BWD, EW_SCALE, EW_AXPY = 0, 1, 2, 3, 4


def _ld(t: torch.Tensor) -> int:
    assert t.dim() == 2 and t.stride(1) == 1, "row-major 2-D matrix expected"
    return t.stride(0)


class DeviceOps:
    def __init__(self, device, precision: str = "fp16"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("the training step runs on a CUDA device (B200); there is no CPU path")
        if precision not in ("fp16", "bf16"):
            raise ValueError("training precision: 'fp16' or 'bf16' operands (fp32 master weights, accumulators and gradients)")
        self.precision = precision
        self.lp_dtype = torch.float16 if precision == "fp16" else torch.bfloat16
        self.kind = LP_FP16 if precision == "fp16" else LP_BF16
        self._scratch = None
        # UNIMM_WGRAD_STREAM=1: wgrad GEMMs on a second stream, see linear_backward.  Measured +1 % on the power-capped step (115.8 -> 114.7 ms,
        # profiles/r02_v10_wgrad_stream_ab.txt): not worth a second stream's hazards by default
        self.wgrad_stream = os.environ.get("UNIMM_WGRAD_STREAM", "0") != "0"
        self._side = None
        self._lb_scratch = [None, None]      # two scratch slots: the wgrad of call i still reads slot i % 2 while call i + 1 fills the other
        self._lb_done = [None, None]         # event: the wgrad that last read the slot has finished
        self._lb_calls = 0
        self._lb_last = None
        self.pre16 = os.environ.get("UNIMM_PRE16", "1") != "0"
        self.pre16_perm = os.environ.get("UNIMM_PRE16_PERM", "1") != "0"     # FFN-1 forward through the fragment-ordered epilogue (linear, pre_act32)
        self._wperm = {}
        self._amax = {}          # data_ptr of a gradient tensor -> (device cell holding max |x| as float bits, numel); see linear_backward

    # ------------------------------------------------------------------ plumbing
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def empty32(self, *shape):
        return torch.empty(*shape, device=self.device, dtype=torch.float32)

    def zeros32(self, *shape):
        return torch.zeros(*shape, device=self.device, dtype=torch.float32)

    def empty16(self, *shape):
        return torch.empty(*shape, device=self.device, dtype=self.lp_dtype)

    def scratch(self, nbytes: int):
        if self._scratch is None or self._scratch.numel() < nbytes:
            self._scratch = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=self.device)
        return self._scratch

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        return self._side

    def _lb_slot(self, nbytes: int):
        """The scratch slot of the next split linear_backward; the main stream first waits for the wgrad that last read it."""
        i = self._lb_calls & 1
        self._lb_calls += 1
        if self._lb_done[i] is not None:
            torch.cuda.current_stream(self.device).wait_event(self._lb_done[i])
        sc = self._lb_scratch[i]
        if sc is None or sc.numel() < nbytes:
            self._lb_scratch[i] = sc = torch.empty(int(nbytes * 1.1) + 1024, dtype=torch.uint8, device=self.device)   # the old one is idle: waited above
        return sc

    def join_side(self):
        """Everything the second stream was given (wgrad GEMMs) is ordered before whatever the main stream does next."""
        if self._lb_last is not None:
            torch.cuda.current_stream(self.device).wait_event(self._lb_last)
            self._lb_last = None

    # max |x| of a gradient, left on the device by the kernel that produced it, saves linear_backward its own pass over the tensor
    def begin_step(self):
        self._amax.clear()

    def new_amax_cell(self):
        return self.zeros32(1)

    def register_amax(self, t, cell):
        self._amax[t.data_ptr()] = (cell, t.numel())

    # ------------------------------------------------------------------ element-wise / rows
    def to_lp(self, x32):
        x32 = x32.contiguous()
        out = self.empty16(*x32.shape)
        check(lib.unimm_k_cast_lp(ptr(x32), ptr(out), x32.numel(), self.kind, self.stream))
        return out

    def cast_into(self, x32, out16):
        assert x32.is_contiguous() and out16.is_contiguous() and x32.numel() == out16.numel()
        check(lib.unimm_k_cast_lp(ptr(x32), ptr(out16), x32.numel(), self.kind, self.stream))

    def ew(self, op, a, b=None, out=None, alpha=1.0):
        out = a if out is None else out
        assert a.is_contiguous() and out.is_contiguous() and (b is None or b.is_contiguous())
        check(lib.unimm_t_ew(op, a.numel(), ptr(a), ptr(b), ptr(out), float(alpha), self.stream))
        return out

    def mul(self, a, b):
        return self.ew(EW_MUL, a, b, out=self.empty32(*a.shape))

    def relu_backward(self, dy, y):
        return self.ew(EW_RELU_BWD, dy, y, out=dy)

    def gather_rows(self, src32, idx):
        out = self.empty32(idx.numel(), src32.shape[1])
        check(lib.unimm_t_gather_rows(ptr(src32), _ld(src32), ptr(idx), idx.numel(), src32.shape[1], ptr(out), self.stream))
        return out

    def scatter_add_rows(self, src32, idx, dst32):
        check(lib.unimm_t_scatter_add_rows(ptr(src32), ptr(idx), idx.numel(), src32.shape[1], ptr(dst32), _ld(dst32), self.stream))

    # ------------------------------------------------------------------ embeddings
    def embed_text_sum(self, ids, seg, pos, word, pos_emb, type_emb, type_ext, type_vocab):
        rows, H = ids.numel(), word.shape[1]
        out = self.empty32(rows, H)
        check(lib.unimm_t_embed_text_sum(ptr(ids), ptr(seg), ptr(pos), rows, H, word.shape[0], pos_emb.shape[0], type_vocab, type_ext.shape[0],
                                         ptr(word), ptr(pos_emb), ptr(type_emb), ptr(type_ext), ptr(out), None, self.stream))
        return out

    def embed_text_backward(self, dsum, ids, seg, pos, g_word, g_pos, g_type, g_type_ext, type_vocab):
        check(lib.unimm_t_embed_text_backward(ptr(dsum), ptr(ids), ptr(seg), ptr(pos), ids.numel(), dsum.shape[1], type_vocab, ptr(g_word),
                                              ptr(g_pos), ptr(g_type), ptr(g_type_ext), self.stream))

    # ------------------------------------------------------------------ LayerNorm / GELU
    def layernorm(self, x32, gamma, beta, want32=True, want16=True):
        rows, H = x32.shape
        y32 = self.empty32(rows, H) if want32 else None
        y16 = self.empty16(rows, H) if want16 else None
        check(lib.unimm_k_layernorm(ptr(x32), _ld(x32), rows, H, ptr(gamma), ptr(beta), ptr(y32), ptr(y16), self.kind, self.stream))
        return y32, y16

    def layernorm_backward(self, dy, x32, gamma, g_gamma, g_beta):
        """-> dx; the parameter gradients are written to ``g_gamma`` / ``g_beta``."""
        rows, H = x32.shape
        assert dy.is_contiguous() and x32.is_contiguous()
        dx, cell = self.empty32(rows, H), self.empty32(1)
        check(lib.unimm_k_layernorm_backward_amax(ptr(dy), ptr(x32), rows, H, ptr(gamma), ptr(dx), ptr(g_gamma), ptr(g_beta), ptr(cell), self.stream))
        self.register_amax(dx, cell)
        return dx

    def gelu(self, t32, want32=False, want16=True):
        g32 = self.empty32(*t32.shape) if want32 else None
        g16 = self.empty16(*t32.shape) if want16 else None
        check(lib.unimm_t_gelu(ptr(t32), t32.numel(), ptr(g32), ptr(g16), self.kind, self.stream))
        return g32, g16

    def gelu_backward(self, dy, t32):
        cell = self.empty32(1)
        check(lib.unimm_k_gelu_backward_amax(ptr(dy), ptr(t32), dy.numel(), ptr(dy), ptr(cell), self.stream))
        self.register_amax(dy, cell)
        return dy

    # ------------------------------------------------------------------ projections
    def dropout(self, x32, drop, want16=True):
        """nn.Dropout with the counter-based mask of ``drop = (seed, p)`` -> (y32, y16 | None)."""
        seed, p = drop
        y32 = self.empty32(*x32.shape)
        y16 = self.empty16(*x32.shape) if want16 else None
        check(lib.unimm_t_dropout(ptr(x32), x32.numel(), seed, float(p), ptr(y32), ptr(y16), self.kind, self.stream))
        return y32, y16

    def dropout_backward(self, dy32, drop):
        seed, p = drop
        assert dy32.is_contiguous()
        check(lib.unimm_t_dropout(ptr(dy32), dy32.numel(), seed, float(p), ptr(dy32), None, self.kind, self.stream))
        return dy32

    def linear(self, x16, w16, bias, residual=None, act=ACT_NONE, want32=True, want16=False, pre_act32=False, drop=None):
        """y = act(x W^T + b) (+ residual) on tcgen05 -> (y32 | None, y16 | None); ``pre_act32``: y32 = x W^T + b (the activation's input, kept
        for the backward) while y16 = act(...)."""
        M, K = x16.shape
        N = w16.shape[0]
        if pre_act32:
            assert want32 and want16 and residual is None
            if self.pre16 and N % 4 == 0 and drop is None:
                # the pre-activation only ever meets gelu' in the backward (linear_backward, gelu_t): kept as 16-bit values — half the
                # write traffic of this epilogue and half the read traffic of the backward's pass (UNIMM_PRE16=0: fp32)
                t16, y16 = self.empty16(M, N), self.empty16(M, N)
                kind, w = self.kind, w16
                if self.pre16_perm and N % 32 == 0 and K % 8 == 0 and w16.is_contiguous():
                    # the scoring engine's fragment-ordered weights: a thread's TMEM fragment is then 8 consecutive output columns and both
                    # 16-bit outputs leave as 16-byte stores without the shared-memory transpose.  The permuted copy is made per call
                    # (the weights change every step; 4.7 MB for an FFN-1 matrix, a few microseconds)
                    w = self._wperm.get((N, K))
                    if w is None:
                        w = self._wperm[(N, K)] = self.empty16(N, K)
                    check(lib.unimm_k_permute_w(ptr(w16), ptr(w), N, K, 1, self.stream))
                    kind = self.kind | 0x100
                check(lib.unimm_k_gemm_lp(ptr(x16), _ld(x16), ptr(w), K, M, N, K, ptr(bias), None, 0, act | 0x200, ptr(t16), N, ptr(y16), N, 0, 0,
                                          kind, self.stream))
                return t16, y16
            act = act | 0x100
        if drop is not None:          # out = dropout(x W^T + b) + residual: the mask is applied in the GEMM epilogue
            assert act == ACT_NONE and want32 and not want16
            y32 = self.empty32(M, N)
            check(lib.unimm_t_gemm_drop(ptr(x16), _ld(x16), ptr(w16), _ld(w16), M, N, K, ptr(bias), ptr(residual),
                                        _ld(residual) if residual is not None else 0, drop[0], float(drop[1]), ptr(y32), N, self.kind, self.stream))
            return y32, None
        y32 = self.empty32(M, N) if want32 else None
        y16 = self.empty16(M, N) if want16 else None
        check(lib.unimm_k_gemm_lp(ptr(x16), _ld(x16), ptr(w16), _ld(w16), M, N, K, ptr(bias), ptr(residual), _ld(residual) if residual is not None else 0,
                                  act, ptr(y32), N, ptr(y16), N, 0, 0, self.kind, self.stream))
        return y32, y16

    def linear_f32(self, x32, w32, bias, residual=None, act=ACT_NONE):
        """The same in fp32 on the CUDA cores: the handful of tiny projections (poolers, NSP head, the K = 5 location term)."""
        M, K = x32.shape
        N = w32.shape[0]
        y = self.empty32(M, N)
        check(lib.unimm_k_gemm_f32(ptr(x32), _ld(x32), ptr(w32), _ld(w32), M, N, K, ptr(bias), ptr(residual), _ld(residual) if residual is not None else 0,
                                   act, ptr(y), N, self.stream))
        return y

    def linear_backward(self, dy32, x16, w16, g_w, g_b, need_dx=True, dx_accum=None, gelu_t=None, dx_amax=False, drop=None):
        """dgrad / wgrad / bias gradient of y = x W^T + b.  ``g_w`` [N, K] / ``g_b`` [N] receive the parameter gradients; returns dX
        (``dx_accum`` += dY W when given).  ``gelu_t``: the projection feeds the erf GELU and ``dy32`` is the gradient with respect to
        the GELU's output — the derivative at the saved pre-activation ``gelu_t`` is applied on the fly.  ``dx_amax``: the dgrad GEMM
        leaves max |dX| on the device for the next consumer of dX (``register_amax``)."""
        M, N = dy32.shape
        K = x16.shape[1]
        assert dy32.is_contiguous() and g_w.is_contiguous() and tuple(g_w.shape) == (N, K)
        nbytes = lib.unimm_k_linear_backward_scratch(M, N, K)
        split = self.wgrad_stream and need_dx and M >= 4096          # small projections: not worth two launches' bookkeeping
        sc = self._lb_slot(nbytes) if split else self.scratch(nbytes)
        dx = None
        if need_dx:
            dx = dx_accum if dx_accum is not None else self.empty32(M, K)
            assert dx.is_contiguous() and tuple(dx.shape) == (M, K)
        cell, n = self._amax.pop(dy32.data_ptr(), (None, 0))
        if cell is not None and n != dy32.numel():
            cell = None
        out_cell = self.empty32(1) if (dx_amax and need_dx) else None
        args = (ptr(dy32), N, ptr(x16), _ld(x16), ptr(w16), _ld(w16), M, N, K, ptr(dx), 1 if dx_accum is not None else 0, ptr(g_w), ptr(g_b), ptr(cell),
                ptr(gelu_t), ptr(out_cell), drop[0] if drop else 0, float(drop[1]) if drop else 0.0, ptr(sc), nbytes,
                self.kind | (0x100 if gelu_t is not None and gelu_t.dtype != torch.float32 else 0))      # | 0x100: 16-bit pre-activation
        if not split:
            check(lib.unimm_k_linear_backward_acc(*args, self.stream))
        else:
            # The wgrad GEMM's result is needed only when the gradients are consumed (all-reduce / optimizer): it runs on a second stream,
            # after this call's dgrad, beside what the backward does next on the main stream — the LayerNorm backward, the next
            # projection's pass over its dY — which is HBM-bound and leaves the tensor cores idle.  join_side() is the barrier.
            main, side = torch.cuda.current_stream(self.device), self._side_stream()
            check(lib.unimm_k_linear_backward_phase(*args, 1, C.c_void_p(main.cuda_stream)))
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            check(lib.unimm_k_linear_backward_phase(*args, 2, C.c_void_p(side.cuda_stream)))
            done = torch.cuda.Event()
            done.record(side)
            self._lb_done[(self._lb_calls - 1) & 1] = done
            self._lb_last = done
            x16.record_stream(side)                                   # a saved activation: freed by the caller while the wgrad may still read it
        if out_cell is not None:
            self.register_amax(dx, out_cell)
        return dx

    # ------------------------------------------------------------------ attention
    def attention(self, q16, k16, v16, B, heads, D, Sq, Skv, mask_kind, desc=None, key_mask=None, drop=None):
        """-> (context 16-bit [B*Sq, heads*D], lse fp32 [B, heads, Sq])."""
        o = self.empty16(B * Sq, heads * D)
        lse = self.empty32(B, heads, Sq)
        check(lib.unimm_k_attention_lse(ptr(q16), _ld(q16), ptr(k16), _ld(k16), ptr(v16), _ld(v16), ptr(o), heads * D, B, heads, D, Sq, Skv, mask_kind,
                                        ptr(desc), ptr(key_mask), self.kind, ptr(lse), drop[0] if drop else 0, float(drop[1]) if drop else 0.0, self.stream))
        return o, lse

    def attention_backward(self, q16, k16, v16, o16, lse, dO32, B, heads, D, Sq, Skv, mask_kind, desc, key_mask, dq, dk, dv, amax_cell=None,
                           drop=None):
        """dq / dk / dv: fp32 2-D views (column blocks of the projections' gradient matrices) that receive the result; ``amax_cell``
        (``new_amax_cell``) accumulates max |result| for ``register_amax`` on those matrices."""
        assert dO32.is_contiguous()
        nbytes = lib.unimm_k_attention_backward_scratch(B, heads, D, Sq)
        sc = self.scratch(nbytes)
        known, n = self._amax.pop(dO32.data_ptr(), (None, 0))
        if known is not None and n != dO32.numel():
            known = None
        check(lib.unimm_k_attention_backward(ptr(q16), _ld(q16), ptr(k16), _ld(k16), ptr(v16), _ld(v16), ptr(o16), _ld(o16), ptr(dO32), ptr(lse), B,
                                             heads, D, Sq, Skv, mask_kind, ptr(desc), ptr(key_mask), self.kind, ptr(dq), _ld(dq), ptr(dk), _ld(dk),
                                             ptr(dv), _ld(dv), ptr(amax_cell), ptr(known), drop[0] if drop else 0, float(drop[1]) if drop else 0.0, ptr(sc),
                                             nbytes, self.stream))

    # ------------------------------------------------------------------ heads / losses
    def lm_head_loss_backward(self, h16, e16, bias, labels_i32, weight32, grad_scale, g_e, g_bias):
        """Fused vocabulary GEMM + likelihood / unlikelihood loss, forward and backward: -> (dH fp32, log p per row); dE / dbias are
        WRITTEN to ``g_e`` / ``g_bias``."""
        n, K = h16.shape
        V = e16.shape[0]
        nbytes = lib.unimm_k_lm_head_backward_scratch(n, V, K)
        sc = self.scratch(nbytes)
        dH, logp = self.empty32(n, K), self.empty32(n)
        check(lib.unimm_k_lm_head_backward(ptr(h16), _ld(h16), ptr(e16), _ld(e16), n, V, K, ptr(bias), ptr(labels_i32), ptr(weight32), float(grad_scale),
                                           ptr(dH), ptr(g_e), ptr(g_bias), ptr(logp), ptr(sc), nbytes, self.kind, self.stream))
        return dH, logp

    def lm_ul_value(self, logp, weight32, scale):
        out = self.empty32(1)
        check(lib.unimm_t_lm_ul_value(ptr(logp), ptr(weight32), logp.numel(), float(scale), ptr(out), self.stream))
        return out

    def nsp_ce(self, logits32, labels_i64, nsp_weight, grad_scale):
        B = logits32.shape[0]
        assert logits32.is_contiguous() and logits32.shape[1] == 2
        loss, d = self.empty32(1), self.empty32(B, 2)
        check(lib.unimm_t_nsp_ce(ptr(logits32), ptr(labels_i64), B, ptr(nsp_weight), float(grad_scale), ptr(loss), ptr(d), self.stream))
        return loss, d

    def image_kl(self, logits32, C_real, target32, target_row_i32, image_label_i64, grad_scale):
        rows, ld = logits32.shape
        loss, d, acc = self.empty32(1), self.empty32(rows, ld), self.empty32(2)
        check(lib.unimm_t_image_kl(ptr(logits32), ld, ptr(target32), ptr(target_row_i32), ptr(image_label_i64), rows, C_real, float(grad_scale),
                                   ptr(loss), ptr(d), ld, ptr(acc), self.stream))
        return loss, d

    def nsp_prob0(self, logits32):
        p0 = self.empty32(logits32.shape[0])
        check(lib.unimm_t_nsp_prob0(ptr(logits32), logits32.shape[0], ptr(p0), self.stream))
        return p0

    def nsp_prob0_backward(self, logits32, dp0, dlogits_accum):
        check(lib.unimm_t_nsp_prob0_backward(ptr(logits32), ptr(dp0), logits32.shape[0], ptr(dlogits_accum), self.stream))

    def neural_ndcg_backward(self, y_pred, y_true, grad_scale, temperature=1.0, max_iter=50, tol=1e-6):
        """-> (grad_scale * d neuralNDCG_transposed / d y_pred, per-slate ndcg)  (utils/rank_loss.py:518-581)."""
        rows, n = y_pred.shape
        d, ndcg = self.empty32(rows, n), self.empty32(rows)
        cnt = torch.empty(1, dtype=torch.int32, device=self.device)
        check(lib.unimm_neural_ndcg_backward(ptr(y_pred), ptr(y_true), rows, n, temperature, max_iter, tol, float(grad_scale), ptr(d), ptr(ndcg), ptr(cnt),
                                             self.stream))
        return d, ndcg

    # ------------------------------------------------------------------ optimizer
    def adamw(self, p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, correct_bias, inv_grad_scale, p16):
        check(lib.unimm_t_adamw(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay),
                                int(step), 1 if correct_bias else 0, float(inv_grad_scale), ptr(p16), self.kind, self.stream))


class TimedOps:
    """``DeviceOps`` with every call bracketed by CUDA events on the launch stream: ``report()`` -> per-operation totals of one or
    more steps (bench.py --workload train_step --profile-ops).  Measurement aid only; the events serialise nothing on one stream."""

    def __init__(self, ops: DeviceOps):
        self._ops = ops
        self._events = []

    def __getattr__(self, name):
        attr = getattr(self._ops, name)
        if not callable(attr) or name in ("empty32", "zeros32", "empty16", "scratch", "begin_step", "new_amax_cell", "register_amax", "join_side"):
            return attr

        def timed(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self._ops.device))
            out = attr(*a, **kw)
            e1.record(torch.cuda.current_stream(self._ops.device))
            key = name
            if name in ("linear", "linear_backward") and len(a) >= 3:
                key = f"{name} M={a[0].shape[0]} N={a[2].shape[0] if name == 'linear_backward' else a[1].shape[0]} K={a[1].shape[1]}"
            self._events.append((key, e0, e1))
            return out
        return timed

    def report(self):
        torch.cuda.synchronize(self._ops.device)
        agg = {}
        for key, e0, e1 in self._events:
            n, ms = agg.get(key, (0, 0.0))
            agg[key] = (n + 1, ms + e0.elapsed_time(e1))
        self._events.clear()
        return dict(sorted(agg.items(), key=lambda kv: -kv[1][1]))
