"""Stub of addict.Dict (tfmesos/scheduler.py:9): an auto-vivifying attribute
dict.  Test infrastructure only."""


class Dict(dict):
    def __init__(self, *args, **kw):
        super(Dict, self).__init__()
        for a in args:
            for k, v in dict(a).items():
                self[k] = v
        for k, v in kw.items():
            self[k] = v

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        if name not in self:
            self[name] = Dict()
        return self[name]

    def __setattr__(self, name, value):
        self[name] = value
