"""Word graphs built on the device from raw token ids: the GPU counterpart of the reference's host
code `ClassificationInteractions.convert_text` + `_laplacian_normalize` (interactions.py:334-351, 11-18). A fitter that
ships token ids (800 B per evidence) instead of dense float64 adjacencies (80 kB per evidence) feeds the model with
`nodes` as `doc_content_without_padding_evidences` and `adj` as `docs_adj`."""
import torch

from . import _lib


def build_word_graphs(tokens: torch.Tensor, lengths: torch.Tensor, n_slots: int, window: int):
    """tokens (G,T) int64 CUDA, lengths (G,) int -> nodes (G,n_slots) int64, adj (G,n_slots,n_slots) f32, n_nodes (G,) int32."""
    if not tokens.is_cuda:
        raise RuntimeError("get_b200: tokens must be a CUDA tensor (there is no CPU path)")
    lib = _lib.load()
    tokens = tokens.to(torch.int64).contiguous()
    lengths = lengths.to(device=tokens.device, dtype=torch.int32).contiguous()
    G, T = tokens.shape
    nodes = torch.empty((G, n_slots), dtype=torch.int64, device=tokens.device)
    adj = torch.empty((G, n_slots, n_slots), dtype=torch.float32, device=tokens.device)
    n_nodes = torch.empty((G,), dtype=torch.int32, device=tokens.device)
    _lib.check(lib.get_build_word_graphs(tokens.data_ptr(), lengths.data_ptr(), G, T, int(n_slots), int(window),
                                         nodes.data_ptr(), adj.data_ptr(), n_nodes.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "get_build_word_graphs")
    return nodes, adj, n_nodes
