"""CPU (fp64): the algebraic identity behind the next planned FLOP reduction ("fold the text->image co-attention").

For the text rows of one unit the image keys / values are the same 37 rows, so the text half of BertBiAttention + BertBiOutput.dense2
(models/vilbert_dialog.py:670-698, :745-752)

    ctx = softmax((x Wq^T + bq)_h K_h^T / sqrt(d) + mask) V_h           per head h, concatenated
    y   = ctx Wo^T + bo

equals, with the per-unit matrices  G_h = Wq_h^T K_h^T  [768, 37],  g_h = bq_h K_h^T  [37],  Z_h = V_h Wo_h^T  [37, 768]:

    y = sum_h softmax((x G_h + g_h) / sqrt(d) + mask) Z_h + bo

i.e. two GEMMs with inner sizes 768 -> 8*37 -> 768 instead of 768 -> 1024 (-> 37 -> 1024) -> 768 per text row.  This test pins the
identity (bias and mask handling included) against the oracle's own attention + linear; nothing in the product uses it yet."""
import math

import torch

from oracle import vilbert_oracle as vo


def test_folded_text_to_image_coattention_equals_the_reference_form():
    torch.manual_seed(0)
    dt = torch.float64
    H, Hb, heads, R, rows = 768, 1024, 8, 37, 50
    d = Hb // heads
    x = torch.randn(rows, H, dtype=dt)
    img = torch.randn(R, Hb, dtype=dt)
    Wq, bq = torch.randn(Hb, H, dtype=dt) * 0.05, torch.randn(Hb, dtype=dt) * 0.05
    Wk, bk = torch.randn(Hb, Hb, dtype=dt) * 0.05, torch.randn(Hb, dtype=dt) * 0.05
    Wv, bv = torch.randn(Hb, Hb, dtype=dt) * 0.05, torch.randn(Hb, dtype=dt) * 0.05
    Wo, bo = torch.randn(H, Hb, dtype=dt) * 0.05, torch.randn(H, dtype=dt) * 0.05
    mask = torch.ones(R, dtype=dt)
    mask[30:] = 0                                                     # padded regions
    add = vo.additive_mask(mask[None, None, None, :], dt)             # [1,1,1,R]
    K, V = img @ Wk.T + bk, img @ Wv.T + bv                           # once per unit in either form

    # reference form (the oracle's functions)
    q = x @ Wq.T + bq
    ctx = vo.attention(q[None], K[None], V[None], heads, add)[0]
    want = ctx @ Wo.T + bo

    # folded form
    got = bo.expand(rows, H).clone()
    for h in range(heads):
        sl = slice(h * d, (h + 1) * d)
        G = Wq[sl].T @ K[:, sl].T                                     # [H, R]
        g = bq[sl] @ K[:, sl].T                                       # [R]
        Z = V[:, sl] @ Wo[:, sl].T                                    # [R, H]
        p = torch.softmax((x @ G + g) / math.sqrt(d) + add[0, 0, 0], dim=-1)
        got += p @ Z
    assert (got - want).abs().max().item() < 1e-10
    # the saving: MACs per text row
    ref_macs = H * Hb + heads * R * d * 2 + Hb * H
    fold_macs = 2 * H * heads * R
    assert fold_macs < 0.3 * ref_macs
