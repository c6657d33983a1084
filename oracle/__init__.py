"""CPU oracle for the DiffPhore denoising hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this
package; the product (diffphore_b200/, src/) never does.  Parity status: the reference ships no tests or
golden vectors and its third-party operators (e3nn, torch_cluster, torch_scatter, PyG) are not installable
here, so the third-party restatements are pinned only by the checkpoint's serialized Wigner-3j buffers and
instruction-table sizes ("parity unpinned" for those); the reference's OWN model code is pinned by importing
/root/reference/src/models/score_model_phore.py over shims (oracle/ref_shims.py, tools/make_golden.py).
"""
