"""Extract e3nn's serialized Wigner-3j buffers from the shipped DiffPhore checkpoint.

The reference checkpoint (weights/diffphore_calibrated_warmuped_ft/best_ema_inference_epoch_model.pt)
stores e3nn 0.5.1's real-basis Wigner-3j tensors as buffers (`*._compiled_main_left_right._w3j_*`).
They are mathematical constants; we keep them as a tiny .npz so that random-init models, the oracle and
the CUDA constant tables do not need the checkpoint.  Run in the build container only:
    python tools/extract_w3j.py
"""
import numpy as np, torch, sys, os
ck = sys.argv[1] if len(sys.argv) > 1 else \
    '/root/reference/weights/diffphore_calibrated_warmuped_ft/best_ema_inference_epoch_model.pt'
sd = torch.load(ck, map_location='cpu', weights_only=False)
out = {}
for k, v in sd.items():
    if '_w3j_' in k:
        name = 'w3j_' + k.split('_w3j_')[1]
        a = v.numpy().astype(np.float64)
        # snap to the exact algebraic values (fp32 buffers) -> store fp32 as shipped
        if name in out:
            assert np.array_equal(out[name], v.numpy()), k
        out[name] = v.numpy()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for dst in ('oracle/w3j.npz', 'diffphore_b200/data/w3j.npz'):
    np.savez(os.path.join(root, dst), **out)
print(sorted(out))
