import torch
from oracle.thirdparty import cluster as _c


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow="source_to_target"):
    ei = _c.radius_graph(x.detach().cpu().numpy(), r, None if batch is None else batch.cpu().numpy(), loop,
                         max_num_neighbors)
    return torch.from_numpy(ei)
