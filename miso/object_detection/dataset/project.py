"""Project = images + labels (ref:miso/object_detection/dataset/project.py:7-124, the subset the
inference / crop entry points touch)."""
from __future__ import annotations

from typing import Dict

from .image import ImageMetadata


class Label:
    def __init__(self, id_, name, colour):
        self.id, self.name, self.colour = id_, name, colour


class Project:
    def __init__(self):
        self.filename = ""
        self.task_names: Dict[int, str] = {}
        self.image_dict: Dict[str, ImageMetadata] = {}
        self.label_dict: Dict[str, Label] = {}

    @property
    def label_names(self):
        return [l.name for l in self.label_dict.values()]

    def add_label(self, id_, name, colour):
        if name not in self.label_dict:
            self.label_dict[name] = Label(id_, name, colour)

    def add_image(self, image: ImageMetadata):
        self.image_dict[image.id] = image

    def remove_labelled_images(self):
        self.image_dict = {k: v for k, v in self.image_dict.items() if len(v.boxes) == 0}

    def remove_unlabelled_images(self):
        self.image_dict = {k: v for k, v in self.image_dict.items() if len(v.boxes) > 0}

    def label_counts(self):
        counts = {k: 0 for k in self.label_dict}
        for im in self.image_dict.values():
            for b in im.boxes:
                counts[b.label] = counts.get(b.label, 0) + 1
        return counts

    def summary(self):
        print(f"{len(self.image_dict)} images")
        for k, v in self.label_counts().items():
            print(f"  {k}: {v}")
