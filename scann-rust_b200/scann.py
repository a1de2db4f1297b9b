"""Scann / ScannBuilder façade (src/scann.rs:35-432) over the GPU searchers.

Mode selection follows Scann::with_config (scann.rs:88-100): brute_force → BruteForce; tree + hash → TreeAH;
tree only → Partitioned; hash only → Hashed.  On the GPU:
  BruteForce  → BruteForceSearcher (exact)
  TreeAH      → TreeXHybridSearcher with the LUT16 scan + exact reorder (the north-star path; the
                reference's own `search_tree_ah` is the weaker non-residual "variant B")
  Hashed      → flat AsymmetricHasher on the LUT16 path (16 codes per block)
  Partitioned → LeafScanSearcher.search_partitioned: exact distances inside the L closest leaves (scann.rs:215-253)
Both builder spellings exist: `.tree()/.hash()` (the code, scann.rs:395-411) and `.partitioned()/.hashed()`
(README / BASELINE.json north_star).
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional

import numpy as np

from . import capi, indexing, searchers
from .capi import ScannError
from .searchers import DistanceMeasure


class SearchMode(Enum):
    BruteForce = "BruteForce"
    Partitioned = "Partitioned"
    Hashed = "Hashed"
    TreeAH = "TreeAH"


@dataclass
class ScannConfig:
    """ScannConfig (src/config.rs:11-29), the fields the hot path reads."""
    num_neighbors: int = 10
    distance_measure: DistanceMeasure = DistanceMeasure.SquaredL2
    brute_force: bool = False
    num_partitions: Optional[int] = None
    num_partitions_to_search: int = 10
    hash_num_blocks: Optional[int] = None
    reorder_num_candidates: Optional[int] = None


class Scann:
    def __init__(self, dataset, config: ScannConfig, device: int = 0):
        n = int(dataset.shape[0]) if hasattr(dataset, "shape") else len(dataset)
        if n == 0:
            raise ScannError(capi.INVALID_ARGUMENT, "Dataset cannot be empty")  # scann.rs:66-68
        self.config = config
        self.device = device
        self._dataset = dataset
        self.size = n
        self.dimensionality = int(dataset.shape[1])
        self._bf = None
        self._tree = None
        self._leaf = None
        if config.brute_force:
            self.search_mode = SearchMode.BruteForce
        elif config.num_partitions is not None and config.hash_num_blocks is not None:
            self.search_mode = SearchMode.TreeAH
        elif config.num_partitions is not None:
            self.search_mode = SearchMode.Partitioned
        elif config.hash_num_blocks is not None:
            self.search_mode = SearchMode.Hashed
        else:
            self.search_mode = SearchMode.BruteForce
        if self.search_mode == SearchMode.BruteForce:
            self._bf = searchers.BruteForceSearcher(dataset, config.distance_measure, device)
        elif self.search_mode == SearchMode.Partitioned:
            self._init_partitioned()
        else:
            self._init_tree_ah()

    # -- constructors mirroring scann.rs:60-137
    @classmethod
    def with_config(cls, dataset, config: ScannConfig, device: int = 0):
        return cls(dataset, config, device)

    @classmethod
    def brute_force(cls, dataset, device: int = 0):
        return cls(dataset, ScannConfig(brute_force=True), device)

    @classmethod
    def partitioned(cls, dataset, num_partitions: int, partitions_to_search: int, device: int = 0):
        return cls(dataset, ScannConfig(num_partitions=num_partitions, num_partitions_to_search=partitions_to_search),
                   device)

    @classmethod
    def hashed(cls, dataset, num_blocks: int, device: int = 0):
        return cls(dataset, ScannConfig(hash_num_blocks=num_blocks), device)

    def _init_partitioned(self):
        import torch
        x = self._dataset
        if not (type(x).__module__.startswith("torch") and x.is_cuda):
            x = torch.as_tensor(np.ascontiguousarray(x, np.float32)).cuda(self.device)
        K = min(int(self.config.num_partitions), self.size)
        centers = indexing.kmeans(x, K, 20, 7)
        assign = indexing.assign_partitions(x, centers, self.device)
        order = torch.argsort(assign.long(), stable=True)
        off = torch.zeros((centers.shape[0] + 1,), dtype=torch.int64, device=x.device)
        off[1:] = torch.cumsum(torch.bincount(assign.long(), minlength=centers.shape[0]), 0)
        self._leaf = searchers.LeafScanSearcher(centers, order.to(torch.int32), off, x, device=self.device)

    def _init_tree_ah(self):
        import torch
        cfg = self.config
        x = self._dataset
        if not (type(x).__module__.startswith("torch") and x.is_cuda):
            x = torch.as_tensor(np.ascontiguousarray(x, np.float32)).cuda(self.device)
        flat = self.search_mode == SearchMode.Hashed
        K = 1 if flat else min(int(cfg.num_partitions), self.size)
        S = int(cfg.hash_num_blocks)
        if self.dimensionality % S != 0:
            raise ScannError(capi.INVALID_ARGUMENT,
                             f"Dimensionality {self.dimensionality} must be divisible by num_subspaces {S}")
        idx = indexing.build_treeah_index(x, K, S, use_residuals=not flat, device=self.device)
        R = cfg.reorder_num_candidates if cfg.reorder_num_candidates else cfg.num_neighbors
        tcfg = searchers.TreeXHybridConfig(num_partitions=K, partitions_to_search=1 if flat else
                                           int(cfg.num_partitions_to_search), use_residuals=not flat,
                                           pre_reorder_multiplier=1.0, distance_measure=cfg.distance_measure)
        self._tree = searchers.TreeXHybridSearcher(tcfg, self.device).build_from_index(
            idx["centers"], idx["codebook"], idx["packed"], idx["ids"], idx["part_offsets"],
            x if cfg.reorder_num_candidates else None)
        self._pre_reorder = int(R)

    def search_batched(self, queries, k: Optional[int] = None):
        """Scann::search_batched (scann.rs:297-303) → (ids, dists, counts)."""
        k = int(k if k is not None else self.config.num_neighbors)
        if self.search_mode == SearchMode.BruteForce:
            return self._bf.search_batched(queries, k)
        if self.search_mode == SearchMode.Partitioned:
            return self._leaf.search_partitioned(queries, k, int(self.config.num_partitions_to_search),
                                                 self.config.distance_measure)
        return self._tree.search_batched(queries, k, pre_reorder_k=max(self._pre_reorder, k))

    def search(self, query, k: Optional[int] = None):
        ids, dists, counts = self.search_batched(np.asarray(query, np.float32)[None, :], k)
        return searchers.results_to_lists(ids, dists, counts)[0]


class ScannBuilder:
    """ScannBuilder (src/scann.rs:364-432)."""

    def __init__(self):
        self.config = ScannConfig()

    def num_neighbors(self, k: int):
        self.config.num_neighbors = int(k)
        return self

    def distance_measure(self, measure):
        self.config.distance_measure = DistanceMeasure(measure)
        return self

    def brute_force(self):
        self.config.brute_force = True
        return self

    def tree(self, num_partitions: int, partitions_to_search: int):
        self.config.num_partitions = int(num_partitions)
        self.config.num_partitions_to_search = int(partitions_to_search)
        return self

    partitioned = tree

    def hash(self, num_blocks: int):
        self.config.hash_num_blocks = int(num_blocks)
        return self

    hashed = hash

    def reorder(self, num_candidates: int):
        self.config.reorder_num_candidates = int(num_candidates)
        return self

    def build(self, dataset, device: int = 0) -> Scann:
        return Scann.with_config(dataset, self.config, device)
