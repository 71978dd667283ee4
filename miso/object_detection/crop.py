"""crop_objects (ref:miso/object_detection/crop.py:9-33): same output tree and file names; the
pixel gather for all boxes of an image is one mb_crop_plan + mb_crop_gather on the device, one
device-to-host copy hands over all crops of the image, and the files are encoded and written by a
thread pool (the reference encodes them one by one with skimage.io.imsave)."""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

from miso.object_detection.dataset.project import Project


def crop_objects(project: Project, output_dir: str, relative_to=None):
    from PIL import Image
    from miso_b200 import detection
    os.makedirs(output_dir, exist_ok=True)
    output_path = Path(output_dir)
    pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    pending = []

    def save(arr, target):
        Image.fromarray(arr).save(target)

    for image in project.image_dict.values():
        if len(image.boxes) == 0:
            continue
        im = np.asarray(Image.open(image.full_path))
        dev = torch.from_numpy(np.array(im, copy=True)).cuda()       # copy: PIL-backed arrays are read-only
        k = len(image.boxes)
        bounds = torch.tensor([[float(v) for v in b.bounds] for b in image.boxes], dtype=torch.float32, device="cuda")
        scores = torch.ones((1, k), dtype=torch.float32, device="cuda")
        counts = torch.tensor([k], dtype=torch.int32, device="cuda")
        out = detection.filter_and_crop([dev], bounds[None], scores, counts, 0.5, boxes_are_xywh=True)
        crops = out.to_host(1 if im.ndim == 2 else im.shape[2])
        path = Path(image.full_path)
        for (_, idx, _xywh, crop), box in zip(crops, image.boxes):
            if relative_to is not None:
                label_path = output_path / path.relative_to(relative_to).parent / box.label
            elif len(project.task_names) > 0:
                label_path = output_path / f"{image.dataset_id} - {project.task_names[image.dataset_id]}" / box.label
            else:
                label_path = output_path / box.label
            label_path.mkdir(parents=True, exist_ok=True)
            s = box.bounds
            filename = f"{path.stem}_{s[0]:.0f}_{s[1]:.0f}_{s[2]:.0f}_{s[3]:.0f}{path.suffix}"
            if crop.size:
                pending.append(pool.submit(save, crop, os.path.join(str(label_path), filename)))
    for f in pending:
        f.result()                      # re-raise encoder / file-system errors
    pool.shutdown()
