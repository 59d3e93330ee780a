"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the UNet training hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.
The product path (``unet-medical-image-contour-segmentation_b200/``) never imports it and fails
loudly when its CUDA library is missing.
"""
