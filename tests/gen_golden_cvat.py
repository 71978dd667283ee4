"""Generate tests/golden/cvat_shapes.json from the reference's own CVAT classes.

Run in the build container (where /root/reference is importable), in a fresh interpreter:  python tests/gen_golden_cvat.py
It imports ref:miso/object_detection/dataset/cvat/cvat_web_api.py (with `lxml` stubbed: absent here), builds the shapes
of a small Project exactly as CvatTask.add_shapes does (lines 407-422; the task object itself needs a server, so its
loop is followed with the same classes: CvatLabeledShape.minimal + CvatLabeledData.minimal(0, shapes=...).to_json())
and stores the inputs next to the JSON text the reference would PATCH."""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = {
    "labels": {"Coccolith": 11, "Coccosphere": 12},
    "images": [
        {"path": "a.png", "frame": 0, "boxes": [[10.4, 20.5, 30.25, 41.5, "Coccolith"], [0.5, 1.5, 2.0, 2.0, "Coccosphere"]]},
        {"path": "b.png", "frame": 3, "boxes": [[56.82802170936761, 5.0, 17.671978449834047, 30.0, "Coccolith"]]},
        {"path": "c.png", "frame": 4, "boxes": []},
    ],
}


def main():
    lxml = types.ModuleType("lxml")
    lxml.etree = types.ModuleType("lxml.etree")
    sys.modules.setdefault("lxml", lxml)
    sys.modules.setdefault("lxml.etree", lxml.etree)
    sys.path.insert(0, "/root/reference")
    import miso.object_detection.dataset.cvat.cvat_web_api as api
    from miso.object_detection.dataset.annotation import RectangleAnnotation
    from miso.object_detection.dataset.image import ImageMetadata
    from miso.object_detection.dataset.project import Project
    assert api.__file__.startswith("/root/reference")
    project = Project()
    for im in CASE["images"]:
        meta = ImageMetadata(im["path"], "/data", 0, im["frame"])
        for x, y, w, h, label in im["boxes"]:
            meta.boxes.append(RectangleAnnotation(x, y, w, h, label))
        project.add_image(meta)
    shapes = []
    for key, image in project.image_dict.items():
        for box in image.boxes:
            shapes.append(api.CvatLabeledShape.minimal("rectangle", False, list(box.coords_int), image.frame_id,
                                                       CASE["labels"][box.label], 0))
    text = api.CvatLabeledData.minimal(0, shapes=shapes).to_json()
    with open(os.path.join(HERE, "golden", "cvat_shapes.json"), "w") as fh:
        json.dump({"case": CASE, "reference_json": text}, fh, indent=1)
    print(text[:400])


if __name__ == "__main__":
    main()
