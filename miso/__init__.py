"""`miso` entry points kept for drop-in use (python -m miso.cli ...), backed by miso_b200.

Only the post-head detection hot path is re-implemented. The host-side data model
below is the minimum that path needs; the CVAT REST client, training engine and augmentation
transforms of the reference are out of scope and are not provided.
"""
