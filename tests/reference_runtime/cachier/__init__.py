def cachier(*a, **k):
    def deco(fn):
        return fn
    return deco
