"""Sample / plot writers of the reference without scipy.misc: mnist/utils.py:32-67 (save_images, merge,
inverse_transform, imsave) and cifar10/common/misc.py:215-244 (save_images grid).  PNGs are written with zlib directly."""
import struct
import zlib

import numpy as np


def write_png(path, img):
    """img: uint8 [H,W] (grey) or [H,W,3] (RGB)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        h, w = img.shape
        color, rows = 0, img.reshape(h, w)
    else:
        h, w, c = img.shape
        assert c == 3, 'PNG writer handles grey and RGB'
        color, rows = 2, img.reshape(h, w * 3)
    raw = b''.join(b'\x00' + rows[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack('>I', len(data)) + tag + data + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)
    with open(path, 'wb') as f:
        f.write(b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, color, 0, 0, 0)) +
                chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))


def read_png(path):
    """inverse of write_png (8-bit grey / RGB, filter type 0 rows as written above) -- for the round-trip test"""
    data = open(path, 'rb').read()
    assert data[:8] == b'\x89PNG\r\n\x1a\n'
    pos, idat, w = 8, b'', None
    while pos < len(data):
        n, tag = struct.unpack('>I', data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        if tag == b'IHDR':
            w, h, depth, color = struct.unpack('>IIBB', body[:10])
        elif tag == b'IDAT':
            idat += body
        pos += 12 + n
    ch = 3 if color == 2 else 1
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + w * ch)
    assert not rows[:, 0].any()
    img = rows[:, 1:].reshape(h, w, ch)
    return img[..., 0] if ch == 1 else img


def inverse_transform(images):
    """mnist/utils.py:97-98"""
    return (images + 1.) / 2.


def merge(images, size):
    """mnist/utils.py:43-63: [N,h,w,c] -> one (size[0] x size[1]) grid image."""
    h, w = images.shape[1], images.shape[2]
    if images.shape[3] in (3, 4):
        c = images.shape[3]
        img = np.zeros((h * size[0], w * size[1], c))
        for idx, image in enumerate(images):
            i, j = idx % size[1], idx // size[1]
            img[j * h:j * h + h, i * w:i * w + w, :] = image
        return img
    elif images.shape[3] == 1:
        img = np.zeros((h * size[0], w * size[1]))
        for idx, image in enumerate(images):
            i, j = idx % size[1], idx // size[1]
            img[j * h:j * h + h, i * w:i * w + w] = image[:, :, 0]
        return img
    raise ValueError('in merge(images,size) images parameter must have dimensions: HxW or HxWx3 or HxWx4')


def imsave(images, size, path):
    """mnist/utils.py:65-67; scipy.misc.imsave rescales to the image's own [min, max] -> [0, 255] (bytescale)."""
    image = np.squeeze(merge(images, size))
    lo, hi = float(image.min()), float(image.max())
    scale = 255.0 / (hi - lo) if hi > lo else 1.0
    out = np.clip((image - lo) * scale + 0.5, 0, 255).astype(np.uint8)
    write_png(path, out[..., :3] if out.ndim == 3 else out)


def save_images(images, size, image_path):
    """mnist/utils.py:32-33"""
    return imsave(inverse_transform(images), size, image_path)


def save_images_grid(X, save_path):
    """cifar10/common/misc.py:215-244: near-square grid of [N,h,w,3] / [N,h,w] / [N,h*w] samples; floats in [0,1] are scaled by
    255.99, integers are taken as they are."""
    X = np.asarray(X)
    if np.issubdtype(X.dtype, np.floating):
        X = (255.99 * X).astype('uint8')
    n_samples = X.shape[0]
    rows = int(np.sqrt(n_samples))
    while n_samples % rows != 0:
        rows -= 1
    nh, nw = rows, int(n_samples / rows)
    if X.ndim == 2:
        s = int(np.sqrt(X.shape[1]))
        X = np.reshape(X, (X.shape[0], s, s))
    h, w = X[0].shape[:2]
    img = np.zeros((h * nh, w * nw, 3) if X.ndim == 4 else (h * nh, w * nw))
    for n, x in enumerate(X):
        j, i = int(n / nw), int(n % nw)
        img[j * h:j * h + h, i * w:i * w + w] = x
    write_png(save_path, np.clip(img, 0, 255).astype(np.uint8))
