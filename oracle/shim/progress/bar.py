class Bar(object):
    def __init__(self, *a, **k):
        self.suffix = ""

    def next(self):
        pass

    def finish(self):
        pass
