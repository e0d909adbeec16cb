"""128x128 image-folder batch generator with the semantics of the LSUN script's loader (LS/tflib/imagenet.py:8-41, LS =
TG/LSUN_bedrooms; used by LS/wgan_LSUN_Bedrooms128.py:352 for its LSUN-bedrooms folder).  SURVEY.md 8(f) row N4.

`make_generator(path, batch_size)` lists the folder once (the reference shells out to `ls`: names in sorted order) and returns
an epoch function.  Every epoch shuffles THE SAME list again, in place, with `numpy.random.RandomState(epoch_number)` (epoch
numbers start at 1), so the orders compound from epoch to epoch like the reference's (:12-14).  Files are written round-robin
into ONE int32 [batch_size, 3, 128, 128] buffer; greyscale 128x128 images are broadcast to three channels, images of any other
shape are skipped WITHOUT giving up their slot's turn (:20-27: the slot keeps its previous content).  The buffer is yielded
whenever `n > 0 and n % batch_size == 0` -- right after slot 0 received the first file of the next batch, never for the last
partial batch -- and before each yield it is mirrored horizontally with probability 1/2 (numpy's GLOBAL RandomState, :33-35).
The reference mirrors by re-binding its buffer name to a reversed VIEW, so the flips compound and later files are written
through the current orientation; that is reproduced with an orientation flag on one base buffer.  Images are read with
Pillow (the reference's `import Image` is the pre-Pillow module name)."""
import os

import numpy as np


def _read(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.array(im)


def make_generator(path, batch_size):
    epoch_count = [1]
    files = sorted(os.listdir(path))

    def get_epoch():
        base = np.zeros((batch_size, 3, 128, 128), dtype='int32')
        mirrored = False                                  # orientation of the buffer the caller sees relative to `base`
        np.random.RandomState(epoch_count[0]).shuffle(files)
        epoch_count[0] += 1
        for n, name in enumerate(files):
            image = _read(os.path.join(path, name))
            if image.shape == (128, 128):
                image = np.broadcast_to(image, (3, 128, 128))
            elif image.shape == (128, 128, 3):
                image = image.transpose(2, 0, 1)
            else:
                continue
            base[n % batch_size] = image[:, :, ::-1] if mirrored else image
            if n > 0 and n % batch_size == 0:
                if np.random.uniform() > 0.5:             # random horizontal flip of the whole buffer
                    mirrored = not mirrored
                yield (base[:, :, :, ::-1] if mirrored else base,)
    return get_epoch


def load(batch_size, data_dir='/home/crcv/xiang/Tensorflow_DCGAN-master/db/lsun/data'):
    return make_generator(data_dir, batch_size)
