"""ctypes binding of ``libunimm_b200.so`` (declarations mirror ``include/unimm_b200.h`` one to one).

There is no fallback: if the shared library has not been built (``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C unimm_b200/csrc``) importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UNIMM_LIB_PATH") or os.path.join(_HERE, "lib", "libunimm_b200.so")   # override: A/B runs of two builds

PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
LP_BF16, LP_FP16 = 0, 1
MASK_TEXT_SELF, MASK_KEY_VECTOR, MASK_CO_INTERVAL = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "vocab_size", "hidden_size", "num_hidden_layers", "num_attention_heads", "intermediate_size",
        "max_position_embeddings", "type_vocab_size", "v_feature_size", "v_target_size", "v_hidden_size",
        "v_num_hidden_layers", "v_num_attention_heads", "v_intermediate_size", "bi_hidden_size",
        "bi_num_attention_heads", "num_connections")] + [
        ("v_biattention_id", C.c_int32 * 16), ("t_biattention_id", C.c_int32 * 16),
        ("seq_len", C.c_int32), ("num_regions", C.c_int32)]


class SeqDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("ctx", C.c_int32), ("L", C.c_int32), ("last_len", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [
        ("B", C.c_int32),
        ("d_input_ids", C.c_void_p), ("d_token_type_ids", C.c_void_p), ("d_position_ids", C.c_void_p),
        ("d_desc", C.c_void_p), ("d_image_feat", C.c_void_p), ("d_image_loc", C.c_void_p),
        ("d_image_mask", C.c_void_p), ("d_feat_index", C.c_void_p), ("d_masked_lm_labels", C.c_void_p),
        ("d_lm_rows", C.c_void_p), ("n_lm_rows", C.c_int32),
        ("d_lm_weight", C.c_void_p), ("d_next_sentence_label", C.c_void_p), ("d_image_label", C.c_void_p),
        ("d_image_target", C.c_void_p), ("d_nsp_weight", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "d_seq_score", "d_token_logp", "d_token_ul", "d_nsp_scores", "d_losses", "d_sequence_output_t",
        "d_sequence_output_v", "d_prediction_scores_t")]


class HostBatch(C.Structure):
    _fields_ = [("B", C.c_int32), ("U", C.c_int32)] + [(n, C.c_void_p) for n in (
        "h_input_ids", "h_token_type_ids", "h_position_ids", "h_masked_lm_labels", "h_desc", "h_image_feat",
        "h_image_loc", "h_image_mask", "h_feat_index")]


class PackedBatchStruct(C.Structure):
    _fields_ = [
        ("n_units", C.c_int32), ("n_cands", C.c_int32), ("n_text_rows", C.c_int32),
        ("d_input_ids", C.c_void_p), ("d_token_type_ids", C.c_void_p), ("d_position_ids", C.c_void_p), ("d_row_iv", C.c_void_p),
        ("d_image_feat", C.c_void_p), ("d_image_loc", C.c_void_p), ("d_image_mask", C.c_void_p),
        ("d_jobs_text_self", C.c_void_p), ("n_jobs_text_self", C.c_int32), ("max_q_text_self", C.c_int32),
        ("n_jobs_text_ctx", C.c_int32), ("cand_halo", C.c_int32),
        ("d_jobs_t2i", C.c_void_p), ("n_jobs_t2i", C.c_int32), ("max_q_t2i", C.c_int32),
        ("d_jobs_i2t", C.c_void_p), ("n_jobs_i2t", C.c_int32),
        ("d_jobs_img_self", C.c_void_p), ("n_jobs_img_self", C.c_int32),
        ("kv_cap_text", C.c_int32), ("win_cap", C.c_int32),
        ("d_lm_rows", C.c_void_p), ("d_lm_labels", C.c_void_p), ("n_lm_rows", C.c_int32),
        ("d_cand_lm_off", C.c_void_p), ("d_cand_cls_row", C.c_void_p), ("d_cand_img_row", C.c_void_p),
        ("pairs_text_self", C.c_double), ("pairs_i2t", C.c_double), ("n_shared_rows", C.c_int32), ("no_cls_rows", C.c_int32),
        ("d_lm_urows", C.c_void_p), ("d_lm_uidx", C.c_void_p), ("n_lm_unique", C.c_int32),
        ("d_unit_image", C.c_void_p), ("n_images", C.c_int32)]


class ImageBlock(C.Structure):
    """unimm_image_block_t: one image's [rows,S] int64 tensors (reference dataloader layout) and its feature block."""
    _fields_ = [("rows", C.c_int32)] + [(n, C.c_void_p) for n in (
        "input_ids", "token_type_ids", "position_ids", "masked_lm_labels", "desc", "image_feat", "image_loc", "image_mask")]


class FlatBatch(C.Structure):
    _fields_ = [("n_blocks", C.c_int32), ("blocks", C.POINTER(ImageBlock)), ("n_units", C.c_int32),
                ("unit_block", C.c_void_p), ("unit_row0", C.c_void_p), ("unit_rows", C.c_void_p),
                ("scores_only", C.c_int32), ("share_first_mask", C.c_int32), ("verify_shared", C.c_int32)]


# every symbol include/unimm_b200.h declares: name -> (restype, argtypes)
_P, _I, _F = C.c_void_p, C.c_int, C.c_void_p
SYMBOLS = {
    "unimm_last_error": (C.c_char_p, []),
    "unimm_abi_version": (C.c_int, []),
    "unimm_create": (C.c_int, [C.POINTER(Config), _I, _I, _I, C.POINTER(C.c_void_p)]),
    "unimm_destroy": (C.c_int, [_P]),
    "unimm_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "unimm_finalize_weights": (C.c_int, [_P]),
    "unimm_forward": (C.c_int, [_P, C.POINTER(Batch), C.POINTER(Outputs), _P]),
    "unimm_forward_packed": (C.c_int, [_P, C.POINTER(PackedBatchStruct), _P, _P, _P, _P]),
    "unimm_score_packed_host": (C.c_int, [_P, C.POINTER(PackedBatchStruct), _P, _P, _P]),
    "unimm_submit_packed_host": (C.c_int, [_P, C.POINTER(PackedBatchStruct), _I, _P, _P, _P]),
    "unimm_wait_packed": (C.c_int, [_P, _I]),
    "unimm_check_ids": (C.c_int, [_P, _P]),
    "unimm_packer_create": (C.c_int, [_I, _I, _I, _I, C.POINTER(C.c_void_p)]),
    "unimm_packer_destroy": (C.c_int, [_P]),
    "unimm_packer_pack": (C.c_int, [_P, C.POINTER(FlatBatch), _I]),
    "unimm_packer_batch": (C.c_int, [_P, C.POINTER(PackedBatchStruct)]),
    "unimm_packer_desc": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int32)]),
    "unimm_verify_masks": (C.c_int, [_P, _I, _I, _I, _P, _I, _P, _P, _P]),
    "unimm_score_host": (C.c_int, [_P, C.POINTER(HostBatch), _P, _P, _P]),
    "unimm_rank_metrics": (C.c_int, [_P, _I, _I, _P, _P, _P, _P, _P]),
    "unimm_neural_ndcg": (C.c_int, [_P, _P, _I, _I, C.c_float, _I, C.c_float, _P, _P, _P]),
    "unimm_ensemble_normalise": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "unimm_neural_ndcg_backward": (C.c_int, [_P, _P, _I, _I, C.c_float, _I, C.c_float, C.c_float, _P, _P, _P, _P]),
    "unimm_t_nsp_prob0": (C.c_int, [_P, _I, _P, _P]),
    "unimm_t_nsp_prob0_backward": (C.c_int, [_P, _P, _I, _P, _P]),
    "unimm_profile_begin": (C.c_int, [_P]),
    "unimm_profile_end": (C.c_int, [_P, _P, _P, _P, _I]),
    "unimm_profile_bytes": (C.c_int, [_P, _P, _I]),
    "unimm_launch_count": (C.c_int64, []),
    "unimm_reset_launch_count": (None, []),
    "unimm_k_gemm_lp": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P, _I, _I, _I, _I, _P]),
    "unimm_k_permute_w_ln": (C.c_int, [_P, _P, _I, _I, _P]),
    "unimm_k_permute_w": (C.c_int, [_P, _P, _I, _I, _I, _P]),
    "unimm_k_gemm_ln_lp": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P, _P, _P, _I, _P, _I, _I, _P]),
    "unimm_k_gemm_f32": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _I, _P]),
    "unimm_k_lm_head_lp": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I, _P]),
    "unimm_k_lm_head_backward_scratch": (C.c_size_t, [_I, _I, _I]),
    "unimm_k_lm_head_backward": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _P, C.c_float, _P, _P, _P, _P, _P, C.c_size_t, _I, _P]),
    "unimm_k_linear_backward_scratch": (C.c_size_t, [_I, _I, _I]),
    "unimm_k_linear_backward": (C.c_int, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _P, C.c_size_t, _I, _P]),
    "unimm_k_layernorm": (C.c_int, [_P, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "unimm_k_layernorm_backward": (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "unimm_k_gelu_backward": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "unimm_k_cast_lp": (C.c_int, [_P, _P, C.c_int64, _I, _P]),
    "unimm_k_attention_jobs": (C.c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _I, _I, _I, _P]),
    "unimm_k_attention_cross_jobs": (C.c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _I, _I, _P, _I, _I, _I, _P]),
    "unimm_k_attention": (C.c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    # training step
    "unimm_k_linear_backward_acc": (C.c_int, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, C.c_uint32, C.c_float, _P, C.c_size_t, _I,
                                              _P]),
    "unimm_k_linear_backward_phase": (C.c_int, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, C.c_uint32, C.c_float, _P, C.c_size_t, _I,
                                                _I, _P]),
    "unimm_t_dropout": (C.c_int, [_P, C.c_int64, C.c_uint32, C.c_float, _P, _P, _I, _P]),
    "unimm_t_gemm_drop": (C.c_int, [_P, _I, _P, _I, _I, _I, _I, _P, _P, _I, C.c_uint32, C.c_float, _P, _I, _I, _P]),
    "unimm_k_layernorm_backward_amax": (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "unimm_k_gelu_backward_amax": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "unimm_k_attention_lse": (C.c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, C.c_uint32, C.c_float, _P]),
    "unimm_k_attention_backward_scratch": (C.c_size_t, [_I, _I, _I, _I]),
    "unimm_k_attention_backward": (C.c_int, [_P, _I, _P, _I, _P, _I, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P, _I, _P, _I,
                                             _P, _P, C.c_uint32, C.c_float, _P, C.c_size_t, _P]),
    "unimm_t_embed_text_sum": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "unimm_t_embed_text_backward": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "unimm_t_gelu": (C.c_int, [_P, C.c_int64, _P, _P, _I, _P]),
    "unimm_t_lm_ul_value": (C.c_int, [_P, _P, _I, C.c_float, _P, _P]),
    "unimm_t_ew": (C.c_int, [_I, C.c_int64, _P, _P, _P, C.c_float, _P]),
    "unimm_t_gather_rows": (C.c_int, [_P, _I, _P, _I, _I, _P, _P]),
    "unimm_t_scatter_add_rows": (C.c_int, [_P, _P, _I, _I, _P, _I, _P]),
    "unimm_t_nsp_ce": (C.c_int, [_P, _P, _I, _P, C.c_float, _P, _P, _P]),
    "unimm_t_image_kl": (C.c_int, [_P, _I, _P, _P, _P, _I, _I, C.c_float, _P, _P, _I, _P, _P]),
    "unimm_t_adamw": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _I, _I, C.c_float, _P, _I, _P]),
}


class UnimmError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built. Run `make -C unimm_b200/csrc` "
            "(or __graft_entry__.build()). unimm_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != 0:
        raise UnimmError(lib.unimm_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device/host address of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
