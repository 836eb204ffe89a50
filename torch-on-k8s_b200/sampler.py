"""Replica-side data sharding: the same indices as torch.utils.data.DistributedSampler
(torch/utils/data/distributed.py:107-145 — SURVEY.md §8 row a13, "bit-exact on indices"), plus an
in-place re-shard when the peer group is re-formed (elastic add/drop)."""
from __future__ import annotations

import math
from typing import Iterator, List

import torch


class ReplicaSampler(torch.utils.data.Sampler):
    def __init__(self, dataset_len: int, num_replicas: int, rank: int, *, shuffle: bool = True,
                 seed: int = 0, drop_last: bool = False):
        if rank < 0 or rank >= num_replicas:
            raise ValueError("invalid rank %d for %d replicas" % (rank, num_replicas))
        self.n = dataset_len
        self.shuffle = shuffle
        self.seed = seed
        self.drop_last = drop_last
        self.epoch = 0
        self.reform(num_replicas, rank)

    def reform(self, num_replicas: int, rank: int) -> None:
        """Elastic re-form: same permutation, new stride."""
        self.num_replicas = num_replicas
        self.rank = rank
        if self.drop_last and self.n % num_replicas != 0:
            self.num_samples = math.ceil((self.n - num_replicas) / num_replicas)
        else:
            self.num_samples = math.ceil(self.n / num_replicas)
        self.total_size = self.num_samples * num_replicas

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def indices(self) -> List[int]:
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            idx = torch.randperm(self.n, generator=g).tolist()
        else:
            idx = list(range(self.n))
        if not self.drop_last:
            pad = self.total_size - len(idx)
            if pad > 0:
                if pad <= len(idx):
                    idx += idx[:pad]
                else:
                    idx += (idx * math.ceil(pad / len(idx)))[:pad]
        else:
            idx = idx[:self.total_size]
        return idx[self.rank:self.total_size:self.num_replicas]

    def __iter__(self) -> Iterator[int]:
        return iter(self.indices())

    def __len__(self) -> int:
        return self.num_samples
