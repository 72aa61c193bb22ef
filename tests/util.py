"""Shared helpers for the test-suite (golden loading, module construction)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix: str = "") -> list[str]:
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n != "collate_fixture" and n.startswith(prefix)]


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    if "config" in d:
        d["config"] = json.loads(str(d["config"]))
    return d


def params_of(g: dict, dtype=torch.float32, device="cpu") -> dict:
    out = {}
    for k, v in g.items():
        if k.startswith("param."):
            out[k[len("param."):]] = torch.from_numpy(v).to(device=device, dtype=dtype)
    return out


def oracle_forward(g: dict, dtype=torch.float32, requires_grad: bool = False):
    """Run oracle/restatement.py on a golden case's inputs; returns (H_v, params dict)."""
    from oracle import restatement as R

    cfg = g["config"]
    P = params_of(g, dtype)
    if requires_grad:
        for p in P.values():
            p.requires_grad_(True)
    V = torch.from_numpy(g["V"]).to(dtype)
    E = torch.from_numpy(g["E"]).to(dtype)
    if cfg.get("graph_transform"):
        V = (V - torch.from_numpy(g["gt_V_mean"]).to(dtype)) / torch.from_numpy(g["gt_V_scale"]).to(dtype)
        E = (E - torch.from_numpy(g["gt_E_mean"]).to(dtype)) / torch.from_numpy(g["gt_E_scale"]).to(dtype)
    ei = torch.from_numpy(g["edge_index"])
    rev = torch.from_numpy(g["rev_edge_index"])
    V_d = torch.from_numpy(g["V_d"]).to(dtype) if "V_d" in g else None
    H = R.message_passing_forward(
        cfg["kind"], V, E, ei, rev, P["W_i.weight"], P.get("W_i.bias"), P["W_h.weight"], P.get("W_h.bias"),
        P["W_o.weight"], P.get("W_o.bias"), cfg["depth"], cfg.get("activation", "relu"),
        cfg.get("undirected", False), V_d, P.get("W_d.weight"), P.get("W_d.bias"))
    return H, P


def build_engine_module(g: dict, device="cuda", precision="fp32", fused=True):
    """chemprop_b200 module holding the golden case's weights."""
    from chemprop_b200.nn import AtomMessagePassing, BondMessagePassing, GraphTransform, ScaleTransform

    cfg = g["config"]
    gt = None
    if cfg.get("graph_transform"):
        gt = GraphTransform(ScaleTransform(g["gt_V_mean"][0], g["gt_V_scale"][0]),
                            ScaleTransform(g["gt_E_mean"][0], g["gt_E_scale"][0]))
    cls = BondMessagePassing if cfg["kind"] == "bond" else AtomMessagePassing
    mp = cls(d_v=cfg.get("d_v", 72), d_e=cfg.get("d_e", 14), d_h=cfg["d_h"], bias=cfg.get("bias", False),
             depth=cfg["depth"], activation=cfg.get("activation", "relu"), undirected=cfg.get("undirected", False),
             d_vd=cfg.get("d_vd"), graph_transform=gt, precision=precision)
    mp.fused = fused
    mp.load_state_dict({k: v for k, v in params_of(g).items()}, strict=False)
    mp = mp.to(device)
    if gt is not None:
        mp.eval()
    return mp


def golden_bmg(g: dict, device="cuda"):
    from chemprop_b200.data import BatchMolGraph

    bmg = BatchMolGraph.from_tensors(
        torch.from_numpy(g["V"]), torch.from_numpy(g["E"]), torch.from_numpy(g["edge_index"]),
        torch.from_numpy(g["rev_edge_index"]), torch.from_numpy(g["batch"]), int(g["n_mols"]))
    bmg.to(device)
    return bmg
