"""Minimal Distribution base (Pyesian/distributions/Distribution.py): size + sample/store/load."""


class Distribution:
    def __init__(self, size: int):
        self._size = int(size)

    def size(self) -> int:
        return self._size

    def sample(self):
        raise NotImplementedError

    def store(self, path: str):
        raise NotImplementedError

    @classmethod
    def load(cls, path: str):
        raise NotImplementedError
