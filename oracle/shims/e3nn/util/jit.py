def compile_mode(mode):
    def deco(cls):
        return cls
    return deco
