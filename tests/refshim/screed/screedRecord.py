class Record:
    def __init__(self, name=None, sequence=None, quality=None, **kw):
        self.name, self.sequence, self.quality = name, sequence, quality
        self.__dict__.update(kw)

    def __len__(self):
        return len(self.sequence)
