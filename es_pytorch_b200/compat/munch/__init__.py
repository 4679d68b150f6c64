"""Minimal ``munch``: attribute-access dicts (``Munch``, ``munchify``, ``unmunchify``)."""


class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def munchify(x):
    if isinstance(x, dict):
        return Munch((k, munchify(v)) for k, v in x.items())
    if isinstance(x, (list, tuple)):
        return type(x)(munchify(v) for v in x)
    return x


def unmunchify(x):
    if isinstance(x, dict):
        return {k: unmunchify(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(unmunchify(v) for v in x)
    return x
