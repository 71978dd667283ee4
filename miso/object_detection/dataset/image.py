"""Image record (ref:miso/object_detection/dataset/image.py:9-61)."""
from __future__ import annotations

import os
from typing import List

from .annotation import RectangleAnnotation


class ImageMetadata:
    def __init__(self, path, container, dataset_id=0, frame_id=0, metadata=None):
        self.path, self.container = path, container
        self.dataset_id, self.frame_id = dataset_id, frame_id
        self.boxes: List[RectangleAnnotation] = []
        self.metadata = metadata if metadata is not None else {}

    @property
    def id(self):
        return f"{self.dataset_id}_{self.frame_id}_{self.path}"

    @property
    def full_path(self):
        return os.path.join(self.container, self.path)

    @property
    def labels(self):
        return list({b.label for b in self.boxes})
