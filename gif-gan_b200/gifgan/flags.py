"""tf.app.flags stand-in: DEFINE_* with the reference's names/defaults -> argparse namespace (`FLAGS`)."""
import argparse


class Flags:
    def __init__(self):
        self._p = argparse.ArgumentParser()
        self.FLAGS = None

    @staticmethod
    def _bool(v):
        return str(v).lower() in ("1", "true", "t", "yes", "y")

    def DEFINE_integer(self, name, default, help=""):
        self._p.add_argument("--" + name, type=lambda s: int(float(s)), default=default, help=help)

    def DEFINE_float(self, name, default, help=""):
        self._p.add_argument("--" + name, type=float, default=default, help=help)

    def DEFINE_string(self, name, default, help=""):
        self._p.add_argument("--" + name, type=str, default=default, help=help)

    def DEFINE_boolean(self, name, default, help=""):
        self._p.add_argument("--" + name, type=self._bool, nargs="?", const=True, default=default, help=help)

    def add_argument(self, *a, **k):
        self._p.add_argument(*a, **k)

    def parse(self, argv=None):
        self.FLAGS = self._p.parse_args(argv)
        return self.FLAGS
