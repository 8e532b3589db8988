from oracle.thirdparty.scatter import scatter, scatter_mean, scatter_sum  # noqa: F401
