"""The annotation payload the reference uploads after inference, without the REST client:
`CvatTask.add_shapes` (ref:miso/object_detection/dataset/cvat/cvat_web_api.py:407-431) turns every box of a Project
into a rectangle shape — points = RectangleAnnotation.coords_int, frame = the image's frame id, label_id looked up by
label name, group 0 — wraps them as labeled data of version 0 and PATCHes the JSON to
`/tasks/<id>/annotations?action=create`. This module builds exactly that JSON text from a Project whose boxes came
back from the device in one transfer per batch (miso.object_detection.inference); sending it is the caller's business
(the CVAT server and its client are out of scope)."""
from __future__ import annotations

import json
from typing import Dict, Mapping

from miso.object_detection.dataset.project import Project


def shapes_payload(project: Project, label_ids: Mapping[str, int]) -> Dict[str, object]:
    """The body of the annotations PATCH as a dict. `label_ids`: label name -> CVAT label id (the reference reads them
    from the task's metadata, `label_dict_by_name[name]["id"]`). Fields the reference leaves at None (z_order, id,
    outside) are dropped by its serialiser and are therefore not produced."""
    shapes = []
    for image in project.image_dict.values():
        for box in image.boxes:
            if box.label not in label_ids:
                raise KeyError(f"label {box.label!r} has no CVAT id (ref: add_missing_labels runs before add_shapes)")
            shapes.append({
                "type": "rectangle",
                "occluded": False,
                "points": [int(v) for v in box.coords_int],
                "frame": image.frame_id,
                "label_id": label_ids[box.label],
                "group": 0,
                "attributes": [],
            })
    return {"version": 0, "tags": [], "shapes": shapes, "tracks": []}


def shapes_json(project: Project, label_ids: Mapping[str, int]) -> str:
    """Byte for byte what the reference sends: keys sorted, indent 4 (ref: CvatJsonSerializable.to_json)."""
    return json.dumps(shapes_payload(project, label_ids), sort_keys=True, indent=4)
