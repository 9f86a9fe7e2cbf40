"""Helpers of the train / eval loop with the reference's names (kgwas/utils.py:20-45, 181-233)."""
from __future__ import annotations

import os
import pickle
import sys

import numpy as np
import torch
from tqdm import tqdm

from .postprocess import compute_metrics, ldsc_regression_weights  # noqa: F401  (re-exported like the reference)


def print_sys(s):
    print(s, flush=True, file=sys.stderr)


def save_dict(path, obj):
    with open(path, "wb") as f:
        pickle.dump(obj, f)


def load_dict(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def evaluate_minibatch_clean(loader, model, device):
    """kgwas/utils.py:20-39.  (The reference runs this without no_grad and throws the graph away; the
    predictions are identical under no_grad, which is what is used here.)"""
    model.eval()
    pred_all, truth = [], []
    with torch.no_grad():
        for batch in tqdm(loader):
            batch = batch.to(device)
            bs = batch["SNP"].batch_size
            out = model(batch.x_dict, batch.edge_index_dict, bs)
            pred_all.append(out.reshape(-1).detach().cpu().numpy())
            truth.append(batch["SNP"].y[:bs].detach().cpu().numpy())
    return {"pred": np.hstack(pred_all) if pred_all else np.zeros(0), "truth": np.hstack(truth) if truth else np.zeros(0)}


def save_model(model, config, path_dir):
    os.makedirs(path_dir, exist_ok=True)
    torch.save(model.state_dict(), os.path.join(path_dir, "model.pt"))
    save_dict(os.path.join(path_dir, "config.pkl"), config)


def load_pretrained(path, model):
    # lazy (never-used) GAT lin_dst weights are stored as UninitializedParameter objects (as PyG does): allow-list that
    # one class instead of unpickling arbitrary code; KGWAS_UNSAFE_LOAD=1 restores the reference's plain torch.load
    # (kgwas/utils.py:210) for checkpoints that hold anything else
    import torch.nn.parameter as _tp
    file = os.path.join(path, "model.pt")
    try:
        with torch.serialization.safe_globals([_tp.UninitializedParameter, _tp.Parameter]):
            state_dict = torch.load(file, map_location=torch.device("cpu"), weights_only=True)
    except Exception:
        if os.environ.get("KGWAS_UNSAFE_LOAD") != "1":
            raise
        state_dict = torch.load(file, map_location=torch.device("cpu"), weights_only=False)
    if next(iter(state_dict))[:7] == "module.":          # checkpoints written from a DataParallel wrapper
        state_dict = {k[7:]: v for k, v in state_dict.items()}
    model.load_state_dict(state_dict)
    return model
