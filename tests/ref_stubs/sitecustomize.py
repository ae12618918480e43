"""Environment shim for running the reference's UNMODIFIED Python in this image (test infrastructure; picked up automatically
because tests/ref_stubs is first on PYTHONPATH of the subprocesses tests/test_gpu_dropin_live.py starts).

The reference's Blender reader builds its RGB image with `Image.fromarray(np.array(arr*255.0, dtype=np.byte), "RGB")`
(scene/dataset_readers.py:276).  The Pillow it was written against (requirements: no pin, 2023 vintage) took the int8 buffer as
raw bytes; Pillow >= 10.3 (12.2 here) raises "Cannot handle this data type: (1, 1, 3), |i1".  Reading the same bytes as uint8 is
exactly what the old code path did, so that is restored here instead of editing the reference file."""
try:
    import numpy as _np
    from PIL import Image as _Image

    _fromarray = _Image.fromarray

    def fromarray(obj, mode=None):
        a = _np.asarray(obj)
        if a.dtype == _np.int8:
            obj = a.view(_np.uint8)
        return _fromarray(obj, mode)

    _Image.fromarray = fromarray
except ImportError:  # no Pillow / numpy: nothing to shim
    pass
