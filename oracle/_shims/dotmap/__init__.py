"""
TEST INFRASTRUCTURE ONLY.  Minimal ``dotmap.DotMap`` stand-in for importing the reference
renderer (src/render/nerf.py:12,276-316 uses attribute access, nested assignment, .toDict()).
"""


class DotMap(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        if key not in self:
            self[key] = DotMap()
        return self[key]

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, DotMap) else v) for k, v in self.items()}
