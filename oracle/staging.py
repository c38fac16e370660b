"""CPU restatement of the reference's per-sample input preparation (TEST INFRASTRUCTURE: only tests/, smoke() and
bench.py's CPU legs may import this).

    zscore_u8        datasetModelSegwithopencv.__getitem__, model/dataset.py:138-148: after cv2.imread(path, 0) and
                     cv2.resize, ``image = (image - image.mean()) / image.std()`` (numpy: float64, population std) and
                     ``torch.as_tensor(image).float()``
    labels_from_u8   dataset.py:150-157 ``torch.as_tensor(label).long()`` and the trainer's ``y[y != 0] = 1``
                     (model/modelUnet.py:130)

Pinned by tests/golden/staging.npz: outputs of the reference's own dataset class on PNG files written by
tests/golden/make_golden_staging.py (bit-equal to this restatement there)."""
import numpy as np
import torch


def zscore_u8(images: np.ndarray) -> torch.Tensor:
    """images uint8 [N, H, W] (or [N, D, H, W]) -> fp32 tensor [N, 1, ...], each sample normalised on its own"""
    out = []
    for im in np.asarray(images):
        a = im.astype(np.uint8)
        z = (a - a.mean()) / a.std()                     # dataset.py:142, float64
        out.append(torch.as_tensor(z[None]).float())      # :144-148 reshape to (1, H, W), .float()
    return torch.stack(out, 0)


def labels_from_u8(labels: np.ndarray, binarize: bool = True) -> torch.Tensor:
    y = torch.as_tensor(np.asarray(labels)).long()         # dataset.py:157
    if binarize:
        y[y != 0] = 1                                      # modelUnet.py:130
    return y
