from oracle.thirdparty.pyg import (MessagePassing, SchNet, global_add_pool, global_mean_pool,  # noqa: F401
                                   InteractionBlock, CFConv, GaussianSmearing, ShiftedSoftplus)
