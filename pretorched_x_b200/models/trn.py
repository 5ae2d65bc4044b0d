"""Temporal Relation Network heads (reference: pretorched/models/trn.py).

``Relation`` (trn.py:20-56) is ReLU -> Linear(T*F, bottleneck) -> ReLU -> Linear(bottleneck, out) over the
concatenated frame features; ``MultiScaleRelation`` (trn.py:59-113) sums such MLPs over sub-sampled frame
tuples of every scale.  Both keep the reference's parameter names (``relate.1.*``, ``relate.3.*``,
``relations.{i}.relate.*``); the bodies run as tcgen05 GEMMs with the bias / ReLU in the epilogue and the
cross-tuple sum accumulated in fp32 by the second GEMM.

Upstream defects mirrored, not fixed: ``HierarchicalRelation`` with depth > 0
and ``MultiScaleHierarchicalRelation`` raise in the reference (torch.stack shape error), and the ``TRN``
wrapper / ``trn()`` factory cannot be built offline (they need the missing ``pretrainedmodels`` package and a
downloaded backbone).  The degenerate ``HierarchicalRelation`` (depth 0), which is what ``TRN(consensus=
'HTRN')`` actually instantiates, is provided.
"""
import itertools

import numpy as np
import torch
import torch.nn as nn

from .. import engine, ops

__all__ = ['Relation', 'MultiScaleRelation', 'HierarchicalRelation']


def _packed(linear):
    return engine._cached(linear, "pl", engine._sig(linear.weight, linear.bias),
                          lambda: ops.PackedLinear(linear.weight, linear.bias))


class Relation(nn.Module):
    """input[..., num_inputs, in_features] -> output[batch, -1, out_features]"""

    def __init__(self, num_inputs, in_features, out_features, bottleneck_dim=512):
        super().__init__()
        self.num_inputs, self.in_features = num_inputs, in_features
        self.out_features, self.bottleneck_dim = out_features, bottleneck_dim
        self.relate = nn.Sequential(
            nn.ReLU(),
            nn.Linear(num_inputs * in_features, bottleneck_dim),
            nn.ReLU(),
            nn.Linear(bottleneck_dim, out_features),
        )

    def mlp_f16(self, x16, out=None, accumulate=False):
        """x16: fp16 [rows][num_inputs*in_features], already ReLU'd.  Returns fp32 [rows][out_features]."""
        hidden = ops.linear(x16, _packed(self.relate[1]), relu=True)
        return ops.linear(hidden, _packed(self.relate[3]), out_f32=True, out=out, accumulate=accumulate)

    def forward(self, input):
        flat = input.contiguous().view(-1, self.num_inputs * self.in_features)
        x16 = ops.cast_rows(flat, relu=True)          # leading nn.ReLU fused into the fp16 cast
        return self.mlp_f16(x16).view(input.size(0), -1, self.out_features)


class MultiScaleRelation(nn.Module):
    def __init__(self, num_input, in_features, out_features, bottleneck_dim=512, num_relations=3):
        super().__init__()
        self.num_input, self.in_features, self.out_features = num_input, in_features, out_features
        self.num_relations, self.bottleneck_dim = num_relations, bottleneck_dim
        self.scales = list(range(num_input, 1, -1))
        self.relations_scales = [self.return_relationset(num_input, s) for s in self.scales]
        self.subsample_scales = [min(num_relations, len(r)) for r in self.relations_scales]
        self.relations = nn.ModuleList(
            [Relation(s, in_features, out_features, bottleneck_dim) for s in self.scales])

    def return_relationset(self, num_input, num_input_relation):
        return list(itertools.combinations(range(num_input), num_input_relation))

    def sample_tuples(self):
        """The frame tuples one forward pass uses: np.random.choice per scale, exactly as trn.py:103-106
        (so seeding NumPy's global RNG reproduces the reference's choice)."""
        picks = []
        for si in range(len(self.scales)):
            idx = np.random.choice(len(self.relations_scales[si]), self.subsample_scales[si], replace=False)
            picks.append([self.relations_scales[si][i] for i in idx])
        return picks

    def forward(self, input):
        feats = input.contiguous().view(-1, self.num_input, self.in_features)
        rows = feats.shape[0]
        x16 = ops.cast_rows(feats.view(rows, -1), relu=True).view(rows, self.num_input, -1)
        if x16.shape[2] != self.in_features:
            raise ValueError("in_features must be a multiple of 8 for the fp16 frame gather")
        total = torch.zeros((rows, self.out_features), dtype=torch.float32, device=input.device)
        for si, tuples in enumerate(self.sample_tuples()):
            for tup in tuples:
                idx = torch.tensor(tup, dtype=torch.int32, device=input.device)
                gathered = ops.gather_frames(x16, idx)
                self.relations[si].mlp_f16(gathered, out=total, accumulate=True)
        return total.view(input.size(0), -1, self.out_features)


class HierarchicalRelation(nn.Module):
    """Only the depth-0 configuration works upstream (trn.py:116-159); it reduces to one Relation over all
    inputs averaged with nothing else."""

    def __init__(self, num_inputs, in_features, out_features, relation_size=4, relation_dist=1, bottleneck_dim=1024):
        super().__init__()
        self.num_inputs, self.in_features, self.out_features = num_inputs, in_features, out_features
        self.relation_size, self.relation_dist, self.bottleneck_dim = relation_size, relation_dist, bottleneck_dim
        depth = int(np.ceil((num_inputs - relation_size) / (relation_size - 1)))
        if depth > 0:
            raise NotImplementedError(
                "HierarchicalRelation with depth > 0 raises a torch.stack shape error in the reference "
                "(trn.py:155-158); only the degenerate depth-0 form is reproducible")
        self.relations = nn.ModuleList([])
        self.linears = nn.ModuleList([])
        self.final_linear = nn.Linear(in_features, out_features)
        self.final_relation = Relation(num_inputs, in_features, out_features)

    def forward(self, input):
        x = input.view(-1, self.num_inputs, self.in_features)
        return self.final_relation(x)     # torch.stack([out]).mean(0) == out
