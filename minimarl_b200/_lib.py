"""ctypes binding of libminimarl_b200.so -- the only way Python reaches the CUDA kernels.

The argument lists below restate ``include/minimarl_b200.h``; ``tests/test_abi.py`` checks that
every symbol the header declares is exported and bound.  Nothing here touches the GPU at import
time, so the binding also loads on a CPU-only box; compute calls need CUDA tensors.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libminimarl_b200.so")

MM_OK = 0


class CheckersCfg(C.Structure):
    """struct mm_checkers_cfg"""
    _fields_ = [
        ("n_envs", C.c_int32), ("max_steps", C.c_int32), ("full_observable", C.c_int32),
        ("auto_reset", C.c_int32), ("reward", (C.c_float * 3) * 2), ("pos_row", C.c_float * 3),
        ("pos_col", C.c_float * 8),
    ]


class SwitchCfg(C.Structure):
    """struct mm_switch_cfg"""
    _fields_ = [
        ("n_envs", C.c_int32), ("n_agents", C.c_int32), ("max_steps", C.c_int32),
        ("full_observable", C.c_int32), ("clock", C.c_int32), ("step_cost", C.c_float),
        ("goal_reward", C.c_float), ("pos_row", C.c_float * 3), ("pos_col", C.c_float * 7),
    ]


class RolloutIO(C.Structure):
    """struct mm_rollout_io"""
    _fields_ = [
        ("env_state", C.c_void_p), ("obs", C.c_void_p), ("obs_t_stride", C.c_int64),
        ("act_idx", C.c_void_p), ("act_t_stride", C.c_int64), ("act_onehot", C.c_void_p), ("onehot_t_stride", C.c_int64),
        ("rewards", C.c_void_p), ("rew_t_stride", C.c_int64), ("dones", C.c_void_p), ("done_t_stride", C.c_int64),
        ("dones_env", C.c_void_p), ("denv_t_stride", C.c_int64), ("step_offset", C.c_uint64), ("eps", C.c_float),
        ("eps_per_env", C.c_int32), ("do_reset", C.c_int32), ("freeze_after_done", C.c_int32), ("common_reward", C.c_int32),
    ]


_P = C.c_void_p
SIGNATURES = {
    "mm_status_string": (C.c_char_p, [C.c_int]),
    "mm_last_cuda_error": (C.c_char_p, []),
    "mm_version": (C.c_int, []),
    "mm_checkers_reset": (C.c_int, [_P, _P, _P, C.POINTER(CheckersCfg), _P]),
    "mm_checkers_step": (C.c_int, [_P, _P, _P, _P, _P, _P, C.POINTER(CheckersCfg), _P]),
    "mm_checkers_rollout": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_uint64, C.c_uint64,
                                      C.POINTER(CheckersCfg), _P]),
    "mm_checkers_rollout_host": (C.c_int, [_P] * 10 + [C.c_int32, C.c_int32, C.POINTER(CheckersCfg), _P, _P]),
    "mm_checkers_rollout_host_compact": (C.c_int, [_P] * 9 + [C.c_int32, C.c_int32, C.c_int32, C.POINTER(CheckersCfg), _P, _P]),
    "mm_switch_reset": (C.c_int, [_P, _P, _P, C.POINTER(SwitchCfg), _P]),
    "mm_switch_step": (C.c_int, [_P, _P, _P, _P, _P, _P, C.POINTER(SwitchCfg), _P]),
    "mm_switch_rollout": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_uint64, C.c_uint64,
                                    C.POINTER(SwitchCfg), _P]),
    "mm_gae_returns": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int64, C.c_double, C.c_double,
                                 C.c_int32, C.c_int32, C.c_float, C.c_float, _P]),
    "mm_segtree_set": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_double, C.c_int64, C.c_int32, _P]),
    "mm_segtree_reduce": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "mm_segtree_find_prefixsum": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int32, _P]),
    "mm_per_sample": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_double, _P, _P, _P]),
    "mm_per_sample_dev": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, _P, _P]),
    "mm_replay_gather": (C.c_int, [_P] * 13 + [C.c_int32] * 7 + [_P]),
    "mm_replay_gather_stacked": (C.c_int, [_P] * 13 + [C.c_int32] * 7 + [_P]),
    "mm_agentq_param_count": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "mm_agentq_param_offsets": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "mm_agentq_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_int32, _P]),
    "mm_qmixer_param_count": (C.c_int, [C.c_int32, C.c_int32]),
    "mm_qmixer_param_offsets": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "mm_qmixer_forward": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "mm_vdn_mix": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P]),
    "mm_qmixer_tc_param_count": (C.c_int, [C.c_int32, C.c_int32]),
    "mm_qmixer_tc_param_offsets": (C.c_int, [C.c_int32, C.c_int32, _P]),
    "mm_qmixer_tc_forward": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "mm_q_select": (C.c_int, [_P] * 7 + [C.c_int32] * 5 + [_P]),
    "mm_qmix_td": (C.c_int, [_P] * 5 + [C.c_float] * 3 + [C.c_int32] * 2 + [_P] * 6),
    "mm_tc_gemm_selftest": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mm_agentq_tc_param_count": (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    "mm_agentq_tc_param_offsets": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "mm_agentq_tc_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                       C.c_int32, _P]),
    "mm_layernorm_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "mm_layernorm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "mm_gru_seq_fwd": (C.c_int, [_P] * 9 + [C.c_int32, C.c_int64, C.c_int32, _P]),
    "mm_gru_seq_bwd": (C.c_int, [_P] * 11 + [C.c_int32, C.c_int64, C.c_int32, _P]),
    "mm_gru_seq_bwd_scaled": (C.c_int, [_P] * 12 + [C.c_int32, C.c_int64, C.c_int32, _P]),
    "mm_qmix_mixing_fwd": (C.c_int, [_P] * 6 + [C.c_int64, C.c_int32, _P]),
    "mm_qmix_mixing_bwd": (C.c_int, [_P] * 10 + [C.c_int64, C.c_int32, _P]),
    "mm_heaptree_update": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int32, _P]),
    "mm_heaptree_sample": (C.c_int, [_P, C.c_int32, _P, C.c_int32, _P, _P, _P]),
    "mm_heaptree_argmin_leaf": (C.c_int, [_P, C.c_int32, _P, _P]),
    "mm_heaptree_add_batch": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P]),
    "mm_qmixer_tc_pack": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P]),
    "mm_qmixer_bwd_dims": (C.c_int, [C.c_int32, C.POINTER(C.c_int32)]),
    "mm_qmixer_bwd_ctas": (C.c_int, [C.c_int64]),
    "mm_qmixer_bwd_grad_layout": (C.c_int, [C.c_int32, C.POINTER(C.c_int32)]),
    "mm_qmixer_bwd": (C.c_int, [_P] * 8 + [C.c_int32, C.c_int64, C.c_int32, C.c_int32, _P]),
    "mm_qmixer_bwd_unfold": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_int32, _P]),
    "mm_qmixer_bwd_unfold_into": (C.c_int, [_P, _P, C.c_int32, _P, _P, C.c_int32, _P]),
    "mm_ln_head_fwd": (C.c_int, [_P] * 8 + [C.c_int64, C.c_int32, C.c_int32, _P]),
    "mm_ln_head_bwd": (C.c_int, [_P] * 12 + [C.c_int64, C.c_int32, C.c_int32, _P]),
    "mm_gru_wgrad_ctas": (C.c_int, [C.c_int64]),
    "mm_gru_wgrad": (C.c_int, [_P] * 6 + [C.c_int32, _P, _P, C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P]),
    "mm_colsum": (C.c_int, [_P, _P, C.c_int64, C.c_int32, _P]),
    "mm_clip_adam_scratch_doubles": (C.c_int, []),
    "mm_clip_adam_step": (C.c_int, [_P] * 7 + [C.c_int64, C.c_int64, C.c_float, _P] + [C.c_float] * 5 + [_P]),
    "mm_ppo_actor_loss_fwd": (C.c_int, [_P] * 6 + [C.c_int64, C.c_int32, C.c_float, _P, _P, C.c_int64, C.c_int64, _P]),
    "mm_ppo_actor_loss_bwd": (C.c_int, [_P] * 7 + [C.c_int64, C.c_int32, C.c_float, _P, C.c_int64, C.c_int64, _P]),
    "mm_ppo_value_loss_fwd": (C.c_int, [_P] * 5 + [C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_int32, _P, _P, C.c_int64,
                                         C.c_int64, _P]),
    "mm_ppo_value_loss_bwd": (C.c_int, [_P] * 7 + [C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_int32, _P, C.c_int64,
                                         C.c_int64, _P]),
    "mm_ppo_trunk_param_count": (C.c_int, []),
    "mm_ppo_trunk_param_offsets": (C.c_int, [_P]),
    "mm_ppo_trunk_max_ctas": (C.c_int, []),
    "mm_ppo_trunk_set_prof": (C.c_int, [_P]),
    "mm_ppo_trunk_fwd": (C.c_int, [_P, _P, _P, C.c_int64] + [_P] * 5 + [C.c_int32, C.c_int64, C.c_int32, C.c_int32, _P]),
    "mm_ppo_trunk_bwd": (C.c_int, [_P, _P, C.c_int64] + [_P] * 5 + [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _P]),
    "mm_ppo_trunk_dims": (C.c_int, [_P]),
    "mm_tc_lane_probe": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "mm_tc_raw_selftest": (C.c_int, [_P, _P, _P] + [C.c_int32] * 10 + [_P]),
    "mm_ppo_trunk2_dims": (C.c_int, [C.POINTER(C.c_int32)]),
    "mm_ppo_trunk2_set_slots": (C.c_int, [C.c_int32, C.c_int32]),
    "mm_ppo_trunk2_bwd_ctas": (C.c_int, [C.c_int64]),
    "mm_ppo_trunk2_set_debug": (C.c_int, [_P]),
    "mm_ppo_trunk2_scale": (C.c_int, [_P, C.c_int64, _P, _P]),
    "mm_ppo_trunk2_pack": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P]),
    "mm_ppo_trunk2_grad_layout": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "mm_ppo_trunk2_unfold": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mm_ppo_trunk2_unfold_into": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mm_ppo_obs_table": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
    "mm_ppo_trunk2_fwd": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, C.c_int32, _P, _P, _P, C.c_int32, C.c_int64, C.c_int32, _P]),
    "mm_ppo_trunk2_bwd": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int32, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int64,
                                    C.c_int32, _P]),
    "mm_agent_mlp_dims": (C.c_int, [C.POINTER(C.c_int32)]),
    "mm_agent_mlp_bwd_ctas": (C.c_int, [C.c_int64]),
    "mm_agent_mlp_pack": (C.c_int, [_P, _P, C.c_int32, _P]),
    "mm_agent_mlp_grad_layout": (C.c_int, [C.c_int32, C.POINTER(C.c_int32)]),
    "mm_agent_mlp_unfold": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.c_int32, _P]),
    "mm_agent_mlp_unfold_into": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, C.c_int32, _P]),
    "mm_agent_mlp_fwd": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "mm_agent_mlp_bwd": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int64, _P]),
    "mm_tc_raw_f16_selftest": (C.c_int, [_P, _P, _P] + [C.c_int32] * 11 + [_P]),
    "mm_mappo_rollout_param_count": (C.c_int, []),
    "mm_mappo_rollout_param_offsets": (C.c_int, [_P]),
    "mm_mappo_rollout_tc": (C.c_int, [_P] * 11 + [C.c_int32, C.c_int64, C.c_int32, C.c_uint64, C.c_uint64,
                                                  C.POINTER(CheckersCfg), _P]),
    "mm_qmix_rollout_tc": (C.c_int, [_P, _P, _P, _P, C.POINTER(RolloutIO), C.POINTER(CheckersCfg), C.POINTER(C.c_float),
                                     C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, _P]),
    "mm_heaptree_add_set": (C.c_int, [_P, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, C.c_int32, _P]),
    "mm_heaptree_rebuild": (C.c_int, [_P, C.c_int32, _P]),
    "mm_chunk_per_insert_episode": (C.c_int, [_P, C.c_int32, _P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "mm_heaptree_set_leaves": (C.c_int, [_P, C.c_int32, _P, _P, C.c_int32, _P, _P]),
    "mm_rows_scatter": (C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, _P]),
    "mm_toplevel_qnet_param_count": (C.c_int, [C.c_int32, C.c_int32]),
    "mm_toplevel_qnet_save_stride": (C.c_int, []),
    "mm_toplevel_rollout": (C.c_int, [_P, C.c_int32] + [_P] * 11 + [C.c_int32] * 5 + [C.c_float, C.c_uint64, C.c_uint64, _P]),
    "mm_toplevel_qnet_seq": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P] + [C.c_int32] * 5 + [_P]),
    "mm_toplevel_qnet_bwd": (C.c_int, [_P] * 5 + [C.c_int32, _P, _P, _P] + [C.c_int32] * 5 + [_P]),
    "mm_toplevel_td": (C.c_int, [_P] * 7 + [C.c_int32] * 4 + [C.c_float, _P]),
    "mm_vdn_chunk_loss": (C.c_int, [_P] * 10 + [C.c_int32] * 3 + [C.c_float, _P]),
    "mm_toplevel_mix_param_count": (C.c_int, [C.c_int32]),
    "mm_toplevel_mix_save_stride": (C.c_int, []),
    "mm_toplevel_mix_seq": (C.c_int, [_P] * 8 + [C.c_int32] * 4 + [_P]),
    "mm_toplevel_mix_bwd": (C.c_int, [_P] * 11 + [C.c_int32] * 5 + [_P]),
    "mm_qmix_chunk_loss": (C.c_int, [_P] * 9 + [C.c_int32, C.c_int32, C.c_float, _P]),
}

_lib = None


def lib():
    """The loaded library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m minimarl_b200.build` "
                "(minimarl_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status):
    if status != MM_OK:
        l = lib()
        msg = l.mm_status_string(status).decode()
        if status == -2:
            msg += ": " + l.mm_last_cuda_error().decode()
        raise RuntimeError(f"minimarl_b200: {msg}")


def ptr(t):
    """Device pointer of a CUDA tensor (or NULL for None); refuses host tensors."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("minimarl_b200 kernels take CUDA tensors; got a CPU tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("minimarl_b200 kernels take contiguous tensors")
    return t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
