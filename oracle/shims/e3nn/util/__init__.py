from . import codegen, jit  # noqa: F401
