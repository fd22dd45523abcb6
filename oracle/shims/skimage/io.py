"""Minimal stand-in for skimage.io (src/entropy_image_coding.py:5,:61,:105;
src/RDE.py:5): PNG read/write through OpenCV, RGB channel order."""
import cv2 as _cv
import numpy as _np

def imread(fn):
    img = _cv.imread(fn, _cv.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(fn)
    if img.ndim == 3:
        img = _cv.cvtColor(img, _cv.COLOR_BGR2RGB)
    return img

def imsave(fn, img, **kw):
    img = _np.asarray(img)
    if img.ndim == 3:
        img = _cv.cvtColor(img, _cv.COLOR_RGB2BGR)
    if not _cv.imwrite(fn, img):
        raise IOError(fn)
