"""Stub for ``yolo.yolov8_seg_model`` (imported by the reference's train.py:18 and export_model.py:5 but
absent from the reference repository).  Outside the hot path."""


class YOLOv8_Seg_S:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("YOLOv8_Seg_S is not part of the reference repository nor of the B200 hot path")
