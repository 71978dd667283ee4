"""the annotation payload the reference PATCHes to CVAT after inference, built without the client.
tests/golden/cvat_shapes.json holds the JSON text produced by the reference's own classes (tests/gen_golden_cvat.py);
miso.object_detection.dataset.cvat.payload must produce it byte for byte from the same Project."""
import json
import os

import pytest

from miso.object_detection.dataset.annotation import RectangleAnnotation
from miso.object_detection.dataset.cvat.payload import shapes_json, shapes_payload
from miso.object_detection.dataset.image import ImageMetadata
from miso.object_detection.dataset.project import Project

HERE = os.path.dirname(os.path.abspath(__file__))


def load_case():
    with open(os.path.join(HERE, "golden", "cvat_shapes.json")) as fh:
        g = json.load(fh)
    project = Project()
    for im in g["case"]["images"]:
        meta = ImageMetadata(im["path"], "/data", 0, im["frame"])
        for x, y, w, h, label in im["boxes"]:
            meta.boxes.append(RectangleAnnotation(x, y, w, h, label))
        project.add_image(meta)
    return project, g["case"]["labels"], g["reference_json"]


def test_shapes_json_equals_the_reference_text():
    project, labels, want = load_case()
    assert shapes_json(project, labels) == want
    body = shapes_payload(project, labels)
    assert body["version"] == 0 and len(body["shapes"]) == 3 and body["tracks"] == [] and body["tags"] == []
    assert body["shapes"][2]["points"][2] == 75          # x + w summed in float64, rounded half to even, like coords_int


def test_unknown_label_is_an_error():
    project, labels, _ = load_case()
    with pytest.raises(KeyError):
        shapes_payload(project, {"Coccolith": 1})
