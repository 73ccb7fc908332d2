"""Entry point of the image DCGAN (the role of models/recurrent_z/main.py in the reference; options: flags.TABLES["image_gan"])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gifgan import flags, ops  # noqa: E402
from gifgan.model import DCGAN  # noqa: E402
from gifgan.utils import pp  # noqa: E402

# constructor arguments taken verbatim from the options
_PASS_THROUGH = ("image_size", "batch_size", "is_crop", "checkpoint_dir", "sample_dir", "data_dir", "log_dir", "image_glob", "shuffle")


def build(opts):
    """DCGAN for the chosen dataset: MNIST is the conditional 28x28x1 variant (model.py:280-296), everything else RGB."""
    kw = {k: getattr(opts, k) for k in _PASS_THROUGH}
    kw["dataset_name"] = opts.dataset
    if opts.dataset == "mnist":
        kw.update(y_dim=10, output_size=28, c_dim=1)
    else:
        kw.update(output_size=opts.output_size, c_dim=opts.c_dim)
    return DCGAN(None, **kw)


def main(argv=None):
    opts = flags.parse("image_gan", argv)
    pp.pprint(vars(opts))
    for d in (opts.checkpoint_dir, opts.sample_dir):
        os.makedirs(d, exist_ok=True)
    ops.set_precision(opts.precision)
    ops.reset_default_store()
    gan = build(opts)
    if opts.is_train:
        gan.train(opts)
    else:
        gan.load(opts.checkpoint_dir)
    return gan


if __name__ == "__main__":
    main()
