"""64x64 ImageNet batch generators with the reference's semantics (TG/tflib/small_imagenet.py:5-24), for CT_gan_64x64.py
(SURVEY.md 8(f) N4).

`load(batch_size, data_dir)` -> `(train_epoch, valid_epoch)` over `<data_dir>/train_64x64/<i>.png` (1,281,149 files) and
`<data_dir>/valid_64x64/<i>.png` (49,999 files), file names zero-padded to the width of the file count.  One epoch =
one pass over a permutation drawn from `numpy.random.RandomState(epoch_number)` (epoch numbers start at 1, :10-12).
The reference fills ONE int32 [batch_size, 3, 64, 64] buffer round-robin and yields it whenever `n > 0 and n % batch_size
== 0` (:16-17) -- i.e. right AFTER slot 0 has been overwritten by the first file of the next batch, and never for the last
partial batch.  A yielded batch therefore holds files n-batch_size+1 .. n of the permutation, with file n in slot 0.  That
behaviour is reproduced exactly (same buffer object yielded every time, like the reference); images are read with Pillow
(`scipy.misc.imread` no longer exists).
"""
import numpy as np


def _imread(path):
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert('RGB'))


def make_generator(path, n_files, batch_size):
    epoch_count = [1]

    def get_epoch():
        images = np.zeros((batch_size, 3, 64, 64), dtype='int32')
        files = list(range(n_files))
        random_state = np.random.RandomState(epoch_count[0])
        random_state.shuffle(files)
        epoch_count[0] += 1
        width = len(str(n_files))
        for n, i in enumerate(files):
            image = _imread("{}/{}.png".format(path, str(i + 1).zfill(width)))
            images[n % batch_size] = image.transpose(2, 0, 1)
            if n > 0 and n % batch_size == 0:
                yield (images,)
    return get_epoch


def load(batch_size, data_dir='/home/ishaan/data/imagenet64', n_train=1281149, n_valid=49999):
    return (make_generator(data_dir + '/train_64x64', n_train, batch_size),
            make_generator(data_dir + '/valid_64x64', n_valid, batch_size))
