from oracle.thirdparty.e3nn_nn import Activation, BatchNorm, FullyConnectedNet, Gate  # noqa: F401
