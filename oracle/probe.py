"""ORACLE (test infrastructure, not product code): restatement of the in-tree linear probe,
/root/reference/primitive_probing/train.py:14-92 (`LinearEncoder.__init__/forward/compute_loss`),
without the pytorch-lightning / torchmetrics plumbing (absent offline).  BASELINE.json config 1.

Quirks kept on purpose: free_space feeds an already soft-maxed output to
``F.cross_entropy`` (train.py:35 + :78); free_space labels are clamped in place to
``max_forward_steps`` (train.py:65); reachability picks one logit per sample (train.py:72);
localization permutes [B,52,9] -> [B,9,52] before flattening (train.py:70).

Only tests/ and bench.py's cpu_baseline leg may import this file.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

NUM_TARGET_OBJECTS = 52      # len(constants.target_objects), primitive_probing/constants.py:1
MAX_FORWARD_STEPS = 10       # primitive_probing/constants.py:3


class LinearEncoder(nn.Module):
    def __init__(self, embedding_type: str, prediction_type: str):
        super().__init__()
        self.embedding_type, self.prediction_type = embedding_type, prediction_type
        if prediction_type in ("object_presence", "reachability", "free_space"):          # train.py:19-40
            assert embedding_type in ("imagenet_avgpool", "clip_avgpool", "clip_attnpool")
            input_dim = 1024 if embedding_type == "clip_attnpool" else 2048
            if prediction_type == "object_presence":
                output_dim, act = NUM_TARGET_OBJECTS, nn.Sigmoid()
            elif prediction_type == "reachability":
                output_dim, act = 110, nn.Sigmoid()
            else:
                output_dim, act = MAX_FORWARD_STEPS + 1, nn.Softmax(dim=1)
            self.model = nn.Sequential(nn.Linear(input_dim, output_dim), act)
        elif prediction_type == "object_localization":                                    # train.py:42-49
            assert embedding_type in ("imagenet_avgpool", "clip_avgpool")
            self.model = nn.Sequential(nn.AdaptiveAvgPool2d(output_size=(3, 3)),
                                       nn.Conv2d(2048, NUM_TARGET_OBJECTS, kernel_size=1),
                                       nn.Flatten(start_dim=2), nn.Sigmoid())
        else:
            raise NotImplementedError(prediction_type)

    def forward(self, x: torch.Tensor) -> torch.Tensor:                                   # train.py:53-54
        return self.model(x)

    def compute_loss(self, batch, eval: bool = False):                                    # train.py:56-92
        x, y = batch
        pt = self.prediction_type
        if pt == "object_localization":
            y = y.flatten(start_dim=1)
        elif pt == "reachability":
            obj_idx, y = y
            obj_idx = obj_idx.tolist()
        elif pt == "free_space":
            y[y > MAX_FORWARD_STEPS] = MAX_FORWARD_STEPS
        y_pred = self.forward(x)
        if pt == "object_localization":
            y_pred = y_pred.permute(0, 2, 1).flatten(start_dim=1)
        elif pt == "reachability":
            y_pred = y_pred[range(len(obj_idx)), obj_idx]
        if pt in ("object_presence", "object_localization", "reachability"):
            loss = F.binary_cross_entropy(y_pred, y.float())
        else:
            loss = F.cross_entropy(y_pred, y)
        if not eval:
            return loss
        if pt == "reachability":
            acc = ((y_pred > 0.5) == y).float().mean()
        elif pt == "free_space":
            acc = (torch.argmax(y_pred, dim=1) == y).float().mean()
        else:   # torchmetrics F1 (micro, threshold 0.5) over all labels -- train.py:86
            p = (y_pred > 0.5)
            t = y.bool()
            tp = (p & t).sum().float()
            acc = 2 * tp / (p.sum() + t.sum()).clamp(min=1).float()
        return loss, {"accuracy": acc}


def probe_train_step(model: LinearEncoder, optimizer: torch.optim.Optimizer, batch) -> float:
    """One `training_step` + Adam step (train.py:94-97,111-113)."""
    optimizer.zero_grad()
    loss = model.compute_loss(batch)
    loss.backward()
    optimizer.step()
    return float(loss.detach())
