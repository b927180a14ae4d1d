"""The `**kargs` vocabulary of the drop-in boundary.

The reference passes run-time tensors to `Graph_basedSemantiStructure.forward` as keyword arguments whose
names are string constants of `KeyWordSettings` (reference `setting_keywords.py:17-22,26,32,38-47`).
When this package is dropped into the reference tree the reference's own `setting_keywords` is used by
the fitter; the string VALUES below are what actually travels, so only they have to agree.
"""


class KeyWordSettings(object):
    Query_Adj = "query_adj"                     # setting_keywords.py:17
    Evd_Docs_Adj = "docs_adj"                   # :18
    GNN_Window = "gnn_window"                   # :19
    Query_lens = "query_lens"                   # :21
    Doc_lens = "docs_lens"                      # :22
    QueryLensIndices = "query_lens_indices"     # :25
    DocLensIndices = "doc_lens_indices"         # :26
    OutputRankingKey = "output_ranking"         # :32
    QuerySources = "query_sources"              # :38
    DocSources = "doc_sources"                  # :39
    TempLabel = "fc_labels"                     # :40
    DocContentNoPaddingEvidence = "doc_content_without_padding_evidences"      # :41
    QueryContentNoPaddingEvidence = "query_content_without_padding_evidences"  # :42
    EvidenceCountPerQuery = "evd_cnt_each_query"   # :46
    FIXED_NUM_EVIDENCES = "fixed_num_evidences"    # :47
