"""Where does the bf16 mode's error come from?  (test infrastructure: a CPU study over the oracle, not a test and not product code)

The bf16 mode rounds every tensor-core operand once to bf16 (8 significant bits): the activations entering each projection, the
weights, Q / K / V, the softmax probabilities entering P.V.  Everything else (residual stream, LayerNorm, softmax, log-sum-exp) is
fp32.  This script restates exactly that on the CPU by wrapping the oracle's `linear` / `attention` and the LM decoder, with a
switch per operand CLASS, and prints how much each class contributes to the error of the 100 sequence log-likelihoods of one
round of the bench-shape fixture (tests/golden/sweep3x100_*.npz, round 10 = the worst one on the GPU).

    python tests/bf16_rounding_study.py [--perturbed 0|1] [--round 10] [--policies all]

Classes: (stream) x (projection) for weights `w:` and activations `a:`, stream in {t, v, c} (text layers, image layers, connection
layers), projection in {qkv, ao, f1, f2}; `att:` = Q / K / V / P rounding inside the attentions; `lm:` = the LM head (transform
and decoder, both operands).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import vilbert_oracle as vo  # noqa: E402


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def hilo(x):
    """bf16 hi + bf16 lo: what a two-plane operand carries (16 significant bits)."""
    h = bf(x)
    return h + bf(x - h)


def layer_index(name):
    import re
    m = re.search(r"layer\.(\d+)\.", name)
    return int(m.group(1)) if m else -1


def classify(name):
    if name.startswith("cls.predictions"):
        return "lm", "lm"
    if ".c_layer." in name:
        s = "c"
        if any(k in name for k in ("query", "key", "value")):
            p = "qkv"
        elif "biOutput" in name:
            p = "ao"
        elif "intermediate" in name:
            p = "f1"
        else:
            p = "f2"
        return s, p
    if ".v_layer." in name or ".layer." in name:
        s = "v" if ".v_layer." in name else "t"
        if ".self." in name:
            p = "qkv"
        elif "attention.output" in name:
            p = "ao"
        elif "intermediate" in name:
            p = "f1"
        else:
            p = "f2"
        return s, p
    return "o", "o"      # image embedding projections, poolers


def fp16(x):
    return x.to(torch.float16).to(torch.float32)


class RulePolicy:
    """rule(stream, projection, layer, operand in {"w", "a"}) -> "raw" | "bf" | "two" | "fp16"; attention operands: att in the same set."""

    def __init__(self, rule, att="bf"):
        self.rule, self.att_mode = rule, att
        self.att, self.two_att = att != "raw", att == "two"

    @staticmethod
    def _apply(mode, x):
        return {"raw": lambda t: t, "bf": bf, "two": hilo, "fp16": fp16}[mode](x)

    def w(self, key, x, name=""):
        return self._apply(self.rule(key[0], key[1], layer_index(name), "w"), x)

    def a(self, key, x, name=""):
        return self._apply(self.rule(key[0], key[1], layer_index(name), "a"), x)


class Policy:
    """Which classes are rounded to bf16 (`on`) and which of those carry a second plane (`two`)."""

    def __init__(self, on_w, on_a, att, two_w=(), two_a=(), two_att=False):
        self.on_w, self.on_a, self.att, self.two_w, self.two_a, self.two_att = set(on_w), set(on_a), att, set(two_w), set(two_a), two_att

    def w(self, key, x, name=""):
        if key in self.two_w or key[0] + ":*" in self.two_w or "*:" + key[1] in self.two_w:
            return hilo(x)
        return bf(x) if key in self.on_w else x

    def a(self, key, x, name=""):
        if key in self.two_a or key[0] + ":*" in self.two_a or "*:" + key[1] in self.two_a:
            return hilo(x)
        return bf(x) if key in self.on_a else x


ALL = [(s, p) for s in "tvc" for p in ("qkv", "ao", "f1", "f2")] + [("lm", "lm"), ("o", "o")]
_WCACHE = {}


def install(policy):
    def linear(sd, name, x):
        key = classify(name)
        if name not in _WCACHE:                       # the cache is cleared per policy
            _WCACHE[name] = policy.w(key, sd[name + ".weight"].float(), name)
        return torch.nn.functional.linear(policy.a(key, x, name), _WCACHE[name], sd[name + ".bias"].float())

    def attention(q, k, v, heads, add_mask, drop=vo._nodrop, site=""):
        r = (hilo if policy.two_att else bf) if policy.att else (lambda t: t)
        qh, kh, vh = vo._split_heads(r(q), heads), vo._split_heads(r(k), heads), vo._split_heads(r(v), heads)
        scores = torch.matmul(qh, kh.transpose(-1, -2)) / np.sqrt(qh.shape[-1])
        if add_mask is not None:
            scores = scores + add_mask
        probs = torch.softmax(scores, dim=-1)
        return vo._merge_heads(torch.matmul(r(probs), vh))

    def lm_logits(sd, rows):
        h = vo.lm_transform(sd, rows)
        key = ("lm", "lm")
        if "decoder" not in _WCACHE:
            _WCACHE["decoder"] = policy.w(key, sd["cls.predictions.decoder.weight"].float())
        return torch.nn.functional.linear(policy.a(key, h), _WCACHE["decoder"]) + sd["cls.predictions.bias"].float()

    vo.linear, vo.attention, vo.lm_logits = linear, attention, lm_logits


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--perturbed", type=int, default=1)
    ap.add_argument("--round", type=int, default=10)
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--policies", default="classes")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    from conftest import golden_state_dict
    from unimm_b200 import synthetic as syn
    from unimm_b200.config import DEFAULT_CONFIG_PATH, ViLBertConfig
    from unimm_b200.descriptors import dense_co_mask, dense_text_mask
    cfg = ViLBertConfig.from_json_file(DEFAULT_CONFIG_PATH)
    sd = golden_state_dict(cfg, 1 if args.perturbed else 0, bool(args.perturbed))
    (feat, loc, mask), views = syn.synth_dialog_rounds(7, rounds=(1, 5, 10))      # the fixture's image and draw order
    v = views[(1, 5, 10).index(args.round)]
    n = args.n
    d = torch.from_numpy(v.desc[:n])
    R = feat.shape[0]
    batch = {
        "tokens": torch.from_numpy(v.tokens[:n]).long(), "segments": torch.from_numpy(v.segments[:n]).long(),
        "positions": torch.from_numpy(v.positions[:n]).long(), "mask": torch.from_numpy(v.labels[:n]).long(),
        "txt_attention_mask": dense_text_mask(d, 256), "co_attention_mask": dense_co_mask(d, 256).long().unsqueeze(1).repeat(1, R, 1),
        "image_feat": torch.from_numpy(feat).unsqueeze(0).expand(n, -1, -1).contiguous(),
        "image_loc": torch.from_numpy(loc).unsqueeze(0).expand(n, -1, -1).contiguous(),
        "image_mask": torch.from_numpy(mask).unsqueeze(0).expand(n, -1).contiguous(),
    }

    def run(policy):
        _WCACHE.clear()
        install(policy)
        t0 = time.time()
        s = vo.score_candidates(sd, cfg, batch, chunk=50, full_logits=False)[0]
        return s.numpy(), time.time() - t0

    ref, dt = run(Policy((), (), False))
    print(f"fp32 reference: {dt:.1f} s, mean seq log-lik {ref.mean():.3f}")
    g = np.load(os.path.join(ROOT, "tests", "golden", "sweep3x100_perturbed.npz" if args.perturbed else "sweep3x100_default.npz"))
    ridx = list(g["round_ids"]).index(args.round)
    print("oracle vs fixture:", np.abs(ref - g["seq_score"][ridx][:n]).max())

    def report(name, pol):
        s, dt = run(pol)
        e = s - ref
        print(f"{name:40s} max {np.abs(e).max():.3e}  rms {np.sqrt((e ** 2).mean()):.3e}  mean {e.mean():+.3e}   ({dt:.0f} s)", flush=True)
        return e

    report("all bf16", Policy(ALL, ALL, True))
    if args.policies == "classes":
        report("weights only", Policy(ALL, (), False))
        report("activations only", Policy((), ALL, False))
        report("attention operands only", Policy((), (), True))
        for s in ("t", "v", "c", "lm"):
            keys = [k for k in ALL if k[0] == s]
            report(f"stream {s}: w + a", Policy(keys, keys, False))
        for p in ("qkv", "ao", "f1", "f2"):
            keys = [k for k in ALL if k[1] == p]
            report(f"projection {p}: w + a", Policy(keys, keys, False))
    elif args.policies == "fixes":
        lm = [("lm", "lm")]
        report("two-plane W + A: lm", Policy(ALL, ALL, True, two_w=lm, two_a=lm))
        report("two-plane W: lm", Policy(ALL, ALL, True, two_w=lm))
        report("two-plane W everywhere", Policy(ALL, ALL, True, two_w=ALL))
        report("two-plane W + A: lm, c:*", Policy(ALL, ALL, True, two_w=lm + ["c:*"], two_a=lm + ["c:*"]))
        report("two-plane W + A: lm; W: *:f1, *:f2", Policy(ALL, ALL, True, two_w=lm + ["*:f1", "*:f2"], two_a=lm))
        report("two-plane W + A: lm; W: c:*", Policy(ALL, ALL, True, two_w=lm + ["c:*"], two_a=lm))

    elif args.policies == "rules":
        def base(s_, p_, l_, o_):
            return "two" if s_ == "lm" else "bf"
        report("lm two-plane (the shipped bf16 mode)", RulePolicy(base))

        def ln_in_fp16(s_, p_, l_, o_):            # projections whose A operand is a LayerNorm output: both operands fp16
            if s_ == "lm":
                return "two"
            return "fp16" if p_ in ("qkv", "f1") else "bf"
        report("+ LayerNorm-input GEMMs (qkv, f1) in fp16", RulePolicy(ln_in_fp16))

        def t11(s_, p_, l_, o_):
            if s_ == "lm" or (s_ == "t" and l_ == 11 and p_ in ("ao", "f1", "f2")):
                return "two"
            return "bf"
        report("+ last text layer's ao / f1 / f2 two-plane", RulePolicy(t11))

        def t11c5(s_, p_, l_, o_):
            if s_ == "lm" or (s_ == "t" and l_ == 11 and p_ in ("ao", "f1", "f2")) or (s_ == "c" and l_ == 5):
                return "two"
            return "bf"
        report("+ t11 tail and c5 two-plane", RulePolicy(t11c5))

        def f2w(s_, p_, l_, o_):
            if s_ == "lm":
                return "two"
            if p_ in ("qkv", "f1"):
                return "fp16"
            return "two" if (p_ == "f2" and o_ == "w") else "bf"
        report("+ fp16 LN-input GEMMs and two-plane f2 weights", RulePolicy(f2w))


if __name__ == "__main__":
    main()
