"""Host helpers with the reference's names and signatures (src/util/helper_func.py), for callers that build the model
and drive fine-tuning through the reference API:

* `init_parent_model` (helper_func.py:339-385): same arguments, returns (model, parent_states).  Only the MaskRCNN
  architecture -- the one every e-OSVOS config uses (cfgs/meta.yaml:71) -- is on the B200 path; the DeepLab variants
  raise NotImplementedError.
* `early_stopping` (helper_func.py:388-398): stop when the best loss of the last `patience` entries improved on the
  best loss before them by no more than `min_loss_improv`.
* `run_loader` -> `eosvos_b200.util.evaluate.run_frames` (helper_func.py:67-159, MaskRCNN branch).
"""
import numpy as np
import torch


def init_parent_model(architecture, encoder, train_encoder, decoder_norm_layer, replace_batch_with_group_norms,
                      batch_norm, roi_pool_output_sizes, eval_augment_rpn_proposals_mode, box_nms_thresh,
                      maskrcnn_loss, **datasets):
    if architecture != 'MaskRCNN':
        raise NotImplementedError(
            f"architecture {architecture!r}: the B200 path implements the MaskRCNN model of the e-OSVOS configs")
    from ..networks.mask_rcnn import MaskRCNN
    model = MaskRCNN(encoder, num_classes=2, batch_norm=batch_norm, train_encoder=train_encoder,
                     roi_pool_output_sizes=roi_pool_output_sizes,
                     eval_augment_rpn_proposals_mode=eval_augment_rpn_proposals_mode,
                     replace_batch_with_group_norms=replace_batch_with_group_norms, box_nms_thresh=box_nms_thresh,
                     maskrcnn_loss=maskrcnn_loss)
    parent_states = {}
    for name, cfg in datasets.items():
        parent_states[name] = {
            'states': [torch.load(p, map_location='cpu') for p in cfg['paths']],
            'splits': [np.loadtxt(p, dtype=str).tolist() for p in cfg['val_split_files']],
        }
    return model, parent_states


def early_stopping(loss_hist, patience, min_loss_improv):
    if patience is None or len(loss_hist) <= patience:
        return False
    best = min(float(v) for v in loss_hist)
    best_before = min(float(v) for v in loss_hist[:-patience])
    return not abs(best - best_before) > min_loss_improv


NAMED_CONFIGS = {"DAVIS-2017": "meta_davis-2017.yaml", "YouTube-VOS": "meta_youtube-vos.yaml",
                 "e-OSVOS": "eval_e-osvos.yaml", "e-OSVOS-OnA": "eval_e-osvos-OnA.yaml"}


def compose_config(cfg_dir, named=(), overrides=None):
    """The reference's Sacred configuration without Sacred (src/train_meta.py:21-31): `cfgs/meta.yaml` <-
    `cfgs/torch.yaml` <- the named configs in the order given on the command line (`DAVIS-2017`, `YouTube-VOS`,
    `e-OSVOS`, `e-OSVOS-OnA`) <- `key.sub=value` overrides as {"key.sub": value}.  `cfg_dir` = the reference's cfgs/
    directory.  The result is the `_config` dict the `evaluate` / `meta_run` workers take."""
    import copy
    import os
    import yaml

    def merge(d, u):
        for k, v in u.items():
            if isinstance(v, dict) and isinstance(d.get(k), dict):
                merge(d[k], v)
            else:
                d[k] = copy.deepcopy(v)
        return d

    cfg = {}
    for f in ("meta.yaml", "torch.yaml") + tuple(NAMED_CONFIGS[n] for n in named):
        with open(os.path.join(cfg_dir, f)) as fh:
            merge(cfg, yaml.safe_load(fh) or {})
    for key, val in (overrides or {}).items():
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = val
    return cfg
