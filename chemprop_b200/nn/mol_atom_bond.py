"""`MABBondMessagePassing` / `MABAtomMessagePassing` with the module API of
chemprop/nn/message_passing/mol_atom_bond.py: the same depth loop as Bond / AtomMessagePassing, followed by a vertex
read-out (`W_vo`, optional `W_vd` on extra atom descriptors) and an edge read-out (`W_eo` on `[E || H]` per directed
edge, optional `W_ed` on extra bond descriptors); `forward(bmg, V_d=None, E_d=None) -> (H_v | None, H_e | None)`.

Executed on the composed tier (chemprop_b200/composed.py): one libdmpnn kernel per reference op, `tau` / `dropout`
applied by torch in between; per-edge outputs are handed back in the caller's edge order.  CUDA only, like the rest.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import composed
from ..engine import get_layout
from ..exceptions import InvalidShapeError
from .message_passing import DEFAULT_ATOM_FDIM, DEFAULT_BOND_FDIM, DEFAULT_HIDDEN_DIM, get_activation_function


class _MABMessagePassingBase(nn.Module):
    _kind = None

    def __init__(self, d_v: int = DEFAULT_ATOM_FDIM, d_e: int = DEFAULT_BOND_FDIM, d_h: int = DEFAULT_HIDDEN_DIM,
                 bias: bool = False, depth: int = 3, dropout: float = 0.0, activation="relu", undirected: bool = False,
                 d_vd: int | None = None, d_ed: int | None = None, V_d_transform: nn.Module | None = None,
                 E_d_transform: nn.Module | None = None, graph_transform: nn.Module | None = None,
                 return_vertex_embeddings: bool = True, return_edge_embeddings: bool = True):
        super().__init__()
        # same keys as the reference's save_hyperparameters() result (mol_atom_bond.py `__init__`)
        self.hparams = dict(d_v=d_v, d_e=d_e, d_h=d_h, bias=bias, depth=depth, dropout=dropout, activation=activation,
                            undirected=undirected, d_vd=d_vd, d_ed=d_ed, return_vertex_embeddings=return_vertex_embeddings,
                            return_edge_embeddings=return_edge_embeddings, V_d_transform=V_d_transform,
                            E_d_transform=E_d_transform, graph_transform=graph_transform, cls=self.__class__)
        self.return_vertex_embeddings = return_vertex_embeddings
        self.return_edge_embeddings = return_edge_embeddings
        self.W_i, self.W_h, self.W_vo, self.W_vd, self.W_eo, self.W_ed = self.setup(d_v, d_e, d_h, d_vd, d_ed, bias)
        self.depth = depth
        self.undirected = undirected
        self.dropout = nn.Dropout(dropout)
        self.tau = get_activation_function(activation)
        self.V_d_transform = V_d_transform if V_d_transform is not None else nn.Identity()
        self.E_d_transform = E_d_transform if E_d_transform is not None else nn.Identity()
        self.graph_transform = graph_transform if graph_transform is not None else nn.Identity()

    def setup(self, d_v, d_e, d_h, d_vd, d_ed, bias):
        raise NotImplementedError

    def _readouts(self, d_v, d_e, d_h, d_vd, d_ed):
        W_vo = nn.Linear(d_v + d_h, d_h) if self.return_vertex_embeddings else None
        W_eo = nn.Linear(d_e + d_h, d_h) if self.return_edge_embeddings else None
        W_vd = nn.Linear(d_h + d_vd, d_h + d_vd) if d_vd else None
        W_ed = nn.Linear(d_h + d_ed, d_h + d_ed) if d_ed else None
        return W_vo, W_vd, W_eo, W_ed

    @property
    def output_dims(self) -> tuple[int | None, int | None]:
        v = None if not self.return_vertex_embeddings else (self.W_vd or self.W_vo).out_features
        e = None if not self.return_edge_embeddings else (self.W_ed or self.W_eo).out_features
        return v, e

    def _descriptors(self, H: Tensor, X_d: Tensor | None, transform, W_d, W_o, name: str) -> Tensor:
        """Second half of vertex_finalize / edge_finalize: W_d on [H || X_d], dropout, no activation."""
        if X_d is None:
            return H
        X_d = transform(X_d)
        try:
            H = self.dropout(W_d(torch.cat((H.to(W_d.weight.dtype), X_d), dim=1)))
        except (RuntimeError, AttributeError):
            raise InvalidShapeError(name, X_d.shape, [len(H), (W_d.in_features - W_o.out_features) if W_d is not None else 0])
        return H

    def forward(self, bmg, V_d: Tensor | None = None, E_d: Tensor | None = None):
        bmg = self.graph_transform(bmg)
        lay = get_layout(bmg)
        H_v, H_e = composed.mab_forward(self, bmg, lay, self._kind)
        if H_v is not None:
            H_v = self._descriptors(H_v, V_d, self.V_d_transform, self.W_vd, self.W_vo, "V_d")
        if H_e is not None:
            H_e = self._descriptors(H_e, E_d, self.E_d_transform, self.W_ed, self.W_eo, "E_d")
        return H_v, H_e


class MABBondMessagePassing(_MABMessagePassingBase):
    _kind = "bond"

    def setup(self, d_v=DEFAULT_ATOM_FDIM, d_e=DEFAULT_BOND_FDIM, d_h=DEFAULT_HIDDEN_DIM, d_vd=None, d_ed=None, bias=False):
        W_i = nn.Linear(d_v + d_e, d_h, bias)
        W_h = nn.Linear(d_h, d_h, bias)
        W_vo, W_vd, W_eo, W_ed = self._readouts(d_v, d_e, d_h, d_vd, d_ed)
        return W_i, W_h, W_vo, W_vd, W_eo, W_ed


class MABAtomMessagePassing(_MABMessagePassingBase):
    _kind = "atom"

    def setup(self, d_v=DEFAULT_ATOM_FDIM, d_e=DEFAULT_BOND_FDIM, d_h=DEFAULT_HIDDEN_DIM, d_vd=None, d_ed=None, bias=False):
        W_i = nn.Linear(d_v, d_h, bias)
        W_h = nn.Linear(d_e + d_h, d_h, bias)
        W_vo, W_vd, W_eo, W_ed = self._readouts(d_v, d_e, d_h, d_vd, d_ed)
        return W_i, W_h, W_vo, W_vd, W_eo, W_ed
