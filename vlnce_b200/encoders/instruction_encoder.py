"""Instruction encoder parameter container (``instruction_encoder.py:11-94``); the forward
(embedding gather, lengths, input projection, persistent (bi)LSTM or (bi)GRU) runs in the CUDA library
and is sequenced by ``vlnce_b200.engine``."""
from __future__ import annotations

import gzip
import json

import torch
import torch.nn as nn


class InstructionEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        rnn = nn.GRU if config.rnn_type == "GRU" else nn.LSTM  # instruction_encoder.py:27
        self.encoder_rnn = rnn(input_size=config.embedding_size, hidden_size=config.hidden_size,
                               bidirectional=config.bidirectional)
        if config.sensor_uuid == "instruction":
            if config.use_pretrained_embeddings:
                with gzip.open(config.embedding_file, "rt") as f:
                    emb = torch.tensor(json.load(f))
                self.embedding_layer = nn.Embedding.from_pretrained(
                    embeddings=emb, freeze=not config.fine_tune_embeddings)
            else:
                self.embedding_layer = nn.Embedding(num_embeddings=config.vocab_size,
                                                    embedding_dim=config.embedding_size, padding_idx=0)

    @property
    def output_size(self):
        return self.config.hidden_size * (1 + int(self.config.bidirectional))
