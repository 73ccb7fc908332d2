"""Host-side image helpers -- the subset of /root/reference/models/recurrent_z/utils.py that the
train/sample path uses (`transform`, `inverse_transform`, `merge`, `get_image`, `save_images`).
scipy.misc image I/O (removed from SciPy) is replaced by OpenCV."""
from __future__ import division

import pprint

import numpy as np

pp = pprint.PrettyPrinter()


def imread(path, is_grayscale=False):
    import cv2
    if is_grayscale:
        img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    else:
        img = cv2.cvtColor(cv2.imread(path, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB)
    if img is None:
        raise IOError("cannot read image %s" % path)
    return img.astype(np.float64)


def center_crop(x, crop_h, crop_w=None, resize_w=64):
    """utils.py:47-55."""
    import cv2
    if crop_w is None:
        crop_w = crop_h
    h, w = x.shape[:2]
    j = int(round((h - crop_h) / 2.))
    i = int(round((w - crop_w) / 2.))
    return cv2.resize(x[j:j + crop_h, i:i + crop_w], (resize_w, resize_w), interpolation=cv2.INTER_LINEAR)


def transform(image, npx=64, is_crop=True, resize_w=64):
    """utils.py:57-63: x/127.5 - 1."""
    cropped_image = center_crop(image, npx, resize_w=resize_w) if is_crop else image
    return np.array(cropped_image) / 127.5 - 1.


def inverse_transform(images):
    """utils.py:65-66."""
    return (images + 1.) / 2.


def get_image(image_path, image_size, is_crop=True, resize_w=64, is_grayscale=False):
    return transform(imread(image_path, is_grayscale), image_size, is_crop, resize_w)


def merge(images, size):
    """utils.py:35-42: tile [N,h,w,c] into a size[0] x size[1] grid."""
    h, w = images.shape[1], images.shape[2]
    img = np.zeros((h * size[0], w * size[1], 3))
    for idx, image in enumerate(images):
        i = idx % size[1]
        j = idx // size[1]
        img[j * h:j * h + h, i * w:i * w + w, :] = image
    return img


def save_images(images, size, image_path):
    import cv2
    img = merge(inverse_transform(images), size)
    img = np.clip(np.around(img * 255), 0, 255).astype(np.uint8)
    return cv2.imwrite(image_path, cv2.cvtColor(img, cv2.COLOR_RGB2BGR))


def open_video_writer(filename, fps, frame_size):
    """cv2.VideoWriter(filename, 0x20, fps, frame_size) as the reference opens it (z_model_lib.py:303, 0x20 = MPEG-4 in the
    OpenCV 2/3 builds it ran on); OpenCV 4 builds reject that tag for .mp4 ("tag 0x00000020 is not found") and would silently
    write nothing, so fall back to the container's own MPEG-4 tag 'mp4v'."""
    import cv2
    w = cv2.VideoWriter(filename, 0x20, fps, frame_size)
    if not w.isOpened():
        w = cv2.VideoWriter(filename, cv2.VideoWriter_fourcc(*"mp4v"), fps, frame_size)
    if not w.isOpened():
        raise IOError("cannot open a video writer for %s" % filename)
    return w
