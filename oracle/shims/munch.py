"""TEST INFRASTRUCTURE — stand-in for the `munch` package (not installable offline) so that the
UNMODIFIED reference pipeline under oracle/_ref/pipeline imports.  Only munchify() is used
(reference main.py:24-25, utils/render_camera/camera.py:208)."""


class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def munchify(x):
    if isinstance(x, dict):
        return Munch({k: munchify(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return type(x)(munchify(v) for v in x)
    return x
