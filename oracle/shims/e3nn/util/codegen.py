class CodeGenMixin:
    pass
