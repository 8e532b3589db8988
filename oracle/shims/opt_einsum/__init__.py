from oracle.thirdparty.einsum import contract  # noqa: F401
