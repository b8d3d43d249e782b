class BIFReader(object):
    pass


class BIFWriter(object):
    pass
