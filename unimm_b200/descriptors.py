"""Sequence descriptors: the four integers that replace the reference's dense attention masks.

The reference builds, per sequence, a dense ``txt_attention_mask [S,S]`` and ``co_txt_attention_mask
[S]`` (utils/data_utils.py:149-210 generative, :300-354 discriminative) and the model turns them into
fp32 additive tensors re-read by every attention layer (models/vilbert_dialog.py:1396-1431).  Both are
a closed form of ``(mode, ctx, L, last_len)``; the kernels regenerate the mask row by row from those.
This module converts between the two representations at the boundary.
"""
from __future__ import annotations

import torch

GEN, DIS = 0, 1


def dense_text_mask(desc: torch.Tensor, S: int) -> torch.Tensor:
    """[B,S,S] bool mask regenerated from descriptors (same closed form as csrc/common.cuh)."""
    d = desc.long()
    mode, ctx, L, last = (d[:, i].view(-1, 1, 1) for i in range(4))
    T = L + last
    r = torch.arange(S, device=desc.device).view(1, S, 1)
    c = torch.arange(S, device=desc.device).view(1, 1, S)
    gen = ((r == 0) & (c < T)) | ((r >= 1) & (r < ctx) & (c >= 1) & (c < ctx)) \
        | ((r >= ctx) & (r < L) & (c >= 1) & (c <= r)) \
        | ((r >= L) & (r < T) & (((c >= 1) & (c < r - last)) | (c == r)))
    dis = (r < L) & (c < L)
    return torch.where(mode == DIS, dis, gen)


def dense_co_mask(desc: torch.Tensor, S: int) -> torch.Tensor:
    """[B,S] int64 co-attention text mask (columns image queries may attend)."""
    d = desc.long()
    mode, ctx, L = d[:, 0:1], d[:, 1:2], d[:, 2:3]
    c = torch.arange(S, device=desc.device).view(1, S)
    return torch.where(mode == DIS, c < L, (c >= 1) & (c < ctx)).long()


def descriptors_from_masks(attention_mask: torch.Tensor, co_attention_mask: torch.Tensor, verify: bool = True) -> torch.Tensor:
    """Derive int32 [B,4] descriptors from the dense tensors ``VisualDialogEncoder.forward`` receives.

    ``attention_mask`` [B,S,S] (bool or int), ``co_attention_mask`` [B,R,S].  With ``verify`` the dense
    masks are regenerated and compared; any other pattern raises (the CUDA path implements exactly the
    reference encoders' mask family, not arbitrary masks).
    """
    if attention_mask.dim() != 3 or co_attention_mask.dim() != 3:
        raise ValueError("expected attention_mask [B,S,S] and co_attention_mask [B,R,S]")
    B, S, _ = attention_mask.shape
    co = co_attention_mask[:, 0, :].long()
    is_dis = co[:, 0] == 1
    ctx = co.sum(-1) + 1
    # generative rows (utils/data_utils.py:199-209): a visible-copy row i in [ctx, L) allows columns [1, i] — i of them — and the
    # first masked-copy row L allows [1, ctx) and itself — ctx of them.  So L is the first row >= ctx whose allowed-column count
    # is not its own index; this also holds for sequences truncated at S (L + last_len > S, :205-209), where row 0 no longer
    # tells T.  No such row: the visible copy itself runs into S (L >= S): L = S describes the same mask.
    rs = attention_mask.sum(-1, dtype=torch.int64)                       # [B,S] allowed columns per row
    r = torch.arange(S, device=rs.device).view(1, S)
    brk = (r >= ctx.view(-1, 1)) & (rs != r)
    L_gen = torch.where(brk.any(-1), brk.float().argmax(-1), torch.full_like(ctx, S))
    last_gen = L_gen - ctx
    L_dis = rs[:, 0]
    zero = torch.zeros_like(ctx)
    desc = torch.stack([is_dis.long(), torch.where(is_dis, zero, ctx), torch.where(is_dis, L_dis, L_gen),
                        torch.where(is_dis, zero, last_gen)], dim=1).to(torch.int32)
    if verify:
        ok = torch.equal(dense_text_mask(desc, S), attention_mask.bool()) and \
            bool((dense_co_mask(desc, S).unsqueeze(1) == co_attention_mask.long()).all())
        if not ok:
            raise NotImplementedError(
                "attention_mask / co_attention_mask are not of the form produced by the reference's "
                "encode_input_gen / encode_input_dis (truncation at max_seq_len included); "
                "the descriptor-driven kernels do not implement arbitrary dense masks")
    return desc
