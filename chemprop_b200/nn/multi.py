"""`MulticomponentMessagePassing` with the API of chemprop/nn/message_passing/multi.py:13-84: one message-passing
block per component of a multicomponent input (or one block shared by all), each block an engine module."""
from __future__ import annotations

import logging
from typing import Iterable, Sequence

from torch import Tensor, nn

logger = logging.getLogger(__name__)


class MulticomponentMessagePassing(nn.Module):
    def __init__(self, blocks: Sequence[nn.Module], n_components: int | None = None, shared: bool = False):
        super().__init__()
        self.hparams = {"cls": self.__class__, "blocks": [block.hparams for block in blocks],
                        "n_components": n_components, "shared": shared}
        if len(blocks) == 0:
            raise ValueError("arg 'blocks' was empty!")
        if shared and len(blocks) > 1:
            logger.warning("More than 1 block was supplied but 'shared' was True! Using only the 0th block...")
        if shared and n_components is None:
            raise ValueError("'shared' is True, so arg 'n_components' is required!")
        self.n_components = n_components
        self.shared = shared
        self.blocks = nn.ModuleList([blocks[0]] * self.n_components if shared else blocks)

    def __len__(self) -> int:
        return len(self.blocks)

    @property
    def output_dim(self) -> int:
        return sum(block.output_dim for block in self.blocks)

    def forward(self, bmgs: Iterable, V_ds: Iterable[Tensor | None] | None = None) -> list[Tensor]:
        if V_ds is None:
            return [block(bmg) for block, bmg in zip(self.blocks, bmgs)]
        return [block(bmg, V_d) for block, bmg, V_d in zip(self.blocks, bmgs, V_ds)]
