"""BBOX node features on the device -- the per-batch feature step in front of the layers
(/root/reference/src/components/nlp/bbox.py:31-124, called at model_train.py:293 / model_predict.py:127).

The reference loops over every text box in Python (`get_shape`, `get_histogram`) and builds the [n, 13]
matrix on the host for every batch.  Here only the string work stays on the host -- counting letters /
digits / other symbols per box (`str.isalpha`, `str.isdigit` are Unicode-aware host operations) -- and the
arithmetic runs in `gte_bbox_features` (float64 like the original, float32 result), writing straight into
the padded feature matrix the first layer reads."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from . import ops


def text_class_counts(texts: Sequence[str]) -> np.ndarray:
    """int32 [n, 3]: (letters, digits, others) per text with blanks removed (bbox.py:83-90)."""
    out = np.zeros((len(texts), 3), dtype=np.int32)
    for i, t in enumerate(texts):
        lit = num = oth = 0
        for ch in t.replace(" ", ""):
            if ch.isalpha():
                lit += 1
            elif ch.isdigit():
                num += 1
            else:
                oth += 1
        out[i] = (lit, num, oth)
    return out


def bbox_features(boxes, texts_or_counts, device="cuda") -> torch.Tensor:
    """``Bbox.__call__`` for one batch of boxes: boxes [n, 4] ints ([x0, y0, x1, y1]) and either the box texts
    or their pre-computed class counts [n, 3]; returns the [n, 13] fp32 feature matrix on ``device``."""
    boxes = torch.as_tensor(np.asarray(boxes, dtype=np.int32).reshape(-1, 4))
    if len(texts_or_counts) and isinstance(texts_or_counts[0], str):
        counts = text_class_counts(texts_or_counts)
    else:
        counts = np.asarray(texts_or_counts, dtype=np.int32).reshape(-1, 3)
    counts = torch.from_numpy(np.ascontiguousarray(counts))
    dev = torch.device(device)
    return ops.bbox_features(boxes.to(dev), counts.to(dev))
