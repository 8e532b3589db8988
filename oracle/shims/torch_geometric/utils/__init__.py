from oracle.thirdparty.pyg import to_undirected  # noqa: F401
