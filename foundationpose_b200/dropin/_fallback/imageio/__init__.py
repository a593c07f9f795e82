"""Minimal stand-in for imageio (used ONLY when the real package is not installed): RGB-ordered imread / imwrite on
OpenCV, which is all the reference drivers call (datareader.py:112, run_demo.py:79, estimater.py:191,224)."""
import cv2
import numpy as np


def imread(path, *a, **k):
    img = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if img is None:
        raise FileNotFoundError(path)
    if img.ndim == 3 and img.shape[2] == 3:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    elif img.ndim == 3 and img.shape[2] == 4:
        img = cv2.cvtColor(img, cv2.COLOR_BGRA2RGBA)
    return img


def imwrite(path, img, *a, **k):
    img = np.asarray(img)
    if img.ndim == 3 and img.shape[2] == 3:
        img = cv2.cvtColor(img, cv2.COLOR_RGB2BGR)
    elif img.ndim == 3 and img.shape[2] == 4:
        img = cv2.cvtColor(img, cv2.COLOR_RGBA2BGRA)
    if not cv2.imwrite(str(path), img):
        raise IOError(f"could not write {path}")


imsave = imwrite
v2 = v3 = None
