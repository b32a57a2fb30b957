"""Deterministic byte-level tokenizer for the `synthetic:` backbones.

The reference tokenises with the checkpoint's own HF tokenizer
(src/vla_fastvlm/model/fastvlm_adapter.py:119-130, 361-380).  Real checkpoint directories still do
(AutoTokenizer works offline on a local directory); random-init presets have no vocabulary files, so
they use this stand-in, which follows the same call protocol: `tok(list_of_str, padding="longest" |
"max_length", truncation=True, max_length=N, return_tensors="pt")` -> {"input_ids", "attention_mask"},
right- or left-padded according to `padding_side`.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Union

import torch


class SimpleByteTokenizer:
    """utf-8 bytes -> ids in [offset, offset+256); pad id 0.  No special tokens are added."""

    def __init__(self, vocab_size: int, offset: int = 3, padding_side: str = "right") -> None:
        if vocab_size < offset + 256:
            offset = max(1, vocab_size - 256)
        self.vocab_size = int(vocab_size)
        self.offset = int(offset)
        self.pad_token_id = 0
        self.padding_side = padding_side
        self.model_max_length = 8192

    def encode(self, text: str) -> List[int]:
        return [self.offset + (b % max(1, self.vocab_size - self.offset)) for b in text.encode("utf-8")]

    def __call__(self, texts: Union[str, Sequence[str]], padding: Union[bool, str] = "longest",
                 truncation: bool = True, max_length: int = 64, return_tensors: str = "pt") -> Dict[str, torch.Tensor]:
        if isinstance(texts, str):
            texts = [texts]
        rows = [self.encode(t) for t in texts]
        if truncation and max_length is not None:
            rows = [r[:max_length] for r in rows]
        width = max_length if padding == "max_length" else max((len(r) for r in rows), default=0)
        width = max(width, 1) if len(rows) else 0
        ids = torch.full((len(rows), width), self.pad_token_id, dtype=torch.long)
        mask = torch.zeros((len(rows), width), dtype=torch.long)
        for i, r in enumerate(rows):
            n = len(r)
            if n == 0:
                continue
            if self.padding_side == "left":
                ids[i, width - n:] = torch.tensor(r, dtype=torch.long)
                mask[i, width - n:] = 1
            else:
                ids[i, :n] = torch.tensor(r, dtype=torch.long)
                mask[i, :n] = 1
        return {"input_ids": ids, "attention_mask": mask}
