"""Writes a checkpoint directory in the layout DeepSpeed 0.6.7 produces for this model (what src/checkpointing.py:17-22
saves and src/evaluation/evaluate_rl.py:509-511 loads): <dir>/latest, <dir>/<tag>/mp_rank_00_model_states.pt with the
key set of deepspeed/runtime/engine.py:_save_checkpoint (module, buffer_names, optimizer, lr_scheduler,
sparse_tensor_module_names, skipped_steps, global_steps, global_samples, dp_world_size, mp_world_size, ds_config,
ds_version + client state) and, for fp16 runs, an FP16_Optimizer-shaped dict under `optimizer`. DeepSpeed itself is not
installed here (third party, absent from /root/reference): the layout is restated from its published source."""
import os

import torch


def write_deepspeed_style_checkpoint(save_dir, module_state, tag="global_step1234", fp16_optimizer=True, client_state=None):
    path = os.path.join(save_dir, tag)
    os.makedirs(path, exist_ok=True)
    opt = None
    if fp16_optimizer:  # FP16_Optimizer.state_dict() (deepspeed/runtime/fp16/fused_optimizer.py)
        opt = {"dynamic_loss_scale": True, "cur_scale": 32768.0, "cur_iter": 1234, "last_overflow_iter": 1100,
               "scale_factor": 2.0, "scale_window": 1000, "optimizer_state_dict": {"state": {}, "param_groups": []},
               "fp32_groups_flat": [torch.zeros(8)], "clip_grad": 1.0}
    state = {"module": {k: v.clone() for k, v in module_state.items()},
             "buffer_names": ["pos_emb.inv_freq"], "optimizer": opt,
             "lr_scheduler": None, "sparse_tensor_module_names": [], "skipped_steps": 3, "global_steps": 1234,
             "global_samples": 1234 * 64, "dp_world_size": 8, "mp_world_size": 1,
             "ds_config": {"train_micro_batch_size_per_gpu": 4, "fp16": {"enabled": True}}, "ds_version": "0.6.7"}
    state.update(client_state or {"iteration": 1234, "args": {"n_layer": 24}})
    torch.save(state, os.path.join(path, "mp_rank_00_model_states.pt"))
    with open(os.path.join(save_dir, "latest"), "w") as f:
        f.write(tag)
    return path
