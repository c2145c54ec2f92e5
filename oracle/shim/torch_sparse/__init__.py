class SparseTensor(object):
    pass


def cat(*a, **k):
    raise NotImplementedError
