import contextlib


class _Ax(object):
    def set(self, *a, **k):
        return self


@contextlib.contextmanager
def ioff():
    yield


def axvline(*a, **k):
    pass


def savefig(*a, **k):
    pass


def clf(*a, **k):
    pass


def figure(*a, **k):
    return None


def close(*a, **k):
    pass
