"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the lp-gnn hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker (or as the timed CPU
baseline), never as the thing shipped.  The product path (``lp-gnn_b200/``) never
imports this package and fails loudly when its CUDA library is missing.

Contents
--------
``pyg_standins.py``  restatement of the three third-party pieces the reference
                     delegates to (torch_geometric.nn.GraphConv 2.1,
                     torch_sparse.SparseTensor 0.6.12, torch_geometric.utils.to_undirected).
                     They are NOT under /root/reference (pinned only in prose,
                     readme.md:51-52), so their published algorithm is restated.
``ref_import.py``    imports the reference's own arch.py / dataset.py / utils.py / val.py /
                     train.py / scripts/pred_basis.py VERBATIM from /root/reference with the stand-ins injected.  Works only
                     where /root/reference exists (the build container); used to validate
                     ``port.py`` and to generate ``tests/golden/*.npz``.
``port.py``          standalone CPU restatement (numpy / torch-CPU) of the hot path that
                     travels to the GPU box: graph construction, GraphConvTwoDirection,
                     GCN_FC.forward, add_knowledge, inference_gnn, the balanced loss,
                     scaling + cvt_to_features.
``make_golden.py``   generates the committed golden vectors from the verbatim reference.

Pinning status: the reference ships no tests, golden vectors, checkpoints or data for
this path, so parity is pinned against OUTPUTS OF THE REFERENCE
ITSELF run in the build container (ref_import.py -> tests/golden/), with the three
third-party stand-ins cross-checked against an independent dense float64 model.
"""
