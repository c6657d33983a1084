"""CUDA-event timing of the two hot kernels inside the timed region, and the roofline arithmetic of bench.py.

Algorithmic bytes of one dp_tp_scatter launch:  E*(4W + 4 D_in + 4 D_sh + 8) + N_out*4*D_out
Algorithmic flops of one dp_edge_mlp launch:                 2*E*(in*hid + (hid+1)*W)
Algorithmic flops of one dp_conv_fused launch:               E*(2*(in+1)*hid + 2*(hid+1)*W + tp_flops)   (MLP + channel mixing;
    the tensor pipe executes 3 * 112/100 times the first term: 2-way FP16 split, chunks padded 100 -> 112 columns)
"""
import os

import torch


class KernelTimer:
    def __init__(self):
        self.records = []          # (kind, name, ev0, ev1, E (int or pinned 1-elem tensor), meta)

    def start(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def stop(self, kind, name, ev0, n_edges, meta):
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        self.records.append((kind, name, ev0, e1, n_edges, meta))

    def _resolved(self):
        torch.cuda.synchronize()
        for kind, name, e0, e1, n, meta in self.records:
            E = int(n.item()) if torch.is_tensor(n) else int(n)
            yield kind, name, e0.elapsed_time(e1) * 1e-3, E, meta

    def summary(self):
        agg = {}
        for kind, name, sec, E, m in self._resolved():
            a = agg.setdefault(f'{kind}:{name}', dict(launches=0, sec=0.0, bytes=0.0, flops=0.0))
            a['launches'] += 1
            a['sec'] += sec
            if kind == 'tp_scatter':
                a['bytes'] += E * (4 * m['W'] + 4 * m['d_in'] + 4 * m['d_sh'] + 8) + m['n_out'] * 4 * m['d_out']
            elif kind == 'conv_fused':
                a['flops'] += float(E) * (2.0 * (m['in_dim'] + 1) * m['hid'] + 2.0 * (m['hid'] + 1) * m['W'] + m['tp_flops'])
                a['bytes'] += float(E) * (240 + 4 * m['d_in'] + 44) + m['n_out'] * 4 * m['d_out']
            elif kind == 'edge_hidden':
                a['flops'] += 2.0 * E * m['in_dim'] * m['hid']
                a['bytes'] += float(E) * (256 + 240)
            else:
                a['flops'] += 2.0 * E * (m['in_dim'] * m['hid'] + (m['hid'] + 1) * m['W'])
                a['bytes'] += 4.0 * E * m['W']
        out = {}
        for k, a in agg.items():
            out[k] = dict(launches=a['launches'], ms_per_launch=1e3 * a['sec'] / max(a['launches'], 1),
                          gbps=a['bytes'] / a['sec'] / 1e9 if a['sec'] > 0 else None,
                          tflops=a['flops'] / a['sec'] / 1e12 if a['flops'] else None, total_ms=1e3 * a['sec'])
        return out

    def roofline_fused(self, peak_tflops, peak_src, hbm_gbs, dominant='lig3', traffic_bytes=None):
        """Roofline of the dominant kernel = dp_conv_fused of the layer-3 ligand-ligand convolution (`dominant`; the largest
        launch of the step), tensor-pipe bound: the weights never reach HBM.  `all_layers` aggregates every dp_conv_fused
        launch of the timed region.  `equiv_hbm_*` is what the unfused dp_tp_scatter would have had to stream in the same
        time (SURVEY 8d algorithmic bytes); `traffic_bytes` = dram read + write of one ncu --set full capture (per launch)."""
        def agg(select):
            f = s = b = mma = 0.0
            n = 0
            for kind, name, sec, E, m in self._resolved():
                if kind == 'conv_fused' and select(name):
                    f += float(E) * (2.0 * (m['in_dim'] + 1) * m['hid'] + 2.0 * (m['hid'] + 1) * m['W'] + m['tp_flops'])
                    # columns the tensor pipe really multiplies per edge (recorded per launch by Engine._conv: W/100 chunks of 112 for
                    # the path-aligned layout, consecutive 112- or 96-column chunks with the last one trimmed to a multiple of 16 for
                    # the flat-trim layouts of the two kernel generations) + the 64 hidden columns, K = 64, three MMAs per product
                    cols = m.get('mma_cols', m['W'] * 1.12)
                    mma += float(E) * 2.0 * 64 * (cols + 64) * 3
                    b += float(E) * (4 * m['W'] + 4 * m['d_in'] + 4 * 9 + 8) + m['n_out'] * 4 * m['d_out']
                    s += sec
                    n += 1
            if s == 0:
                return None
            return dict(achieved=f / s / 1e12, peak=peak_tflops, unit='TFLOP/s', frac=f / s / 1e12 / peak_tflops, launches=n,
                        flops_per_launch=f / n, ms_per_launch=1e3 * s / n, issued_mma_tflops=mma / s / 1e12,
                        issued_mma_frac=mma / s / 1e12 / peak_tflops, fp32_parity_bound_frac=f / mma,
                        equiv_hbm_gbs=b / s / 1e9, equiv_hbm_frac=b / s / 1e9 / hbm_gbs)
        dom, allk = agg(lambda nm: nm == dominant), agg(lambda nm: True)
        if dom is None:
            return None
        out = dict(kernel=f'conv_fused_kernel<TpL3> ({dominant}: layer-3 ligand-ligand convolution)', bound='tensor')
        out.update(dom)
        out.update(traffic=traffic_bytes, peak_source=f'{peak_src} dense bf16 cuBLAS throughput (sustained)', all_layers=allk,
                   note='achieved = algorithmic FLOPs E*(2*61*60 + 2*61*W + tp_flops) (both MLP layers + channel mixing) / CUDA-event '
                        'time; the fp32-parity FP16 split issues 3 MMAs per product over K = 64 (61 used) and the flat-trim chunk layout rounds W up to a multiple of 16 columns '
                        '(issued_mma_*; fp32_parity_bound_frac = algorithmic / issued FLOPs = the largest `frac` this fp32-parity scheme can reach); equiv_hbm_* = bytes the unfused dp_tp_scatter would stream (SURVEY 8d) / this '
                        'kernel\'s time; traffic = dram__bytes_read.sum + dram__bytes_write.sum of this launch in '
                        'profiles/ncu_conv_fused_r2.csv')
        return out

    def roofline(self, peak_gbs, peak_src):
        """Aggregate over every dp_tp_scatter launch of the timed region (the kernel BASELINE.json's metric names)."""
        tot_b = tot_s = 0.0
        n = 0
        for kind, name, sec, E, m in self._resolved():
            if kind == 'tp_scatter':
                tot_b += E * (4 * m['W'] + 4 * m['d_in'] + 4 * m['d_sh'] + 8) + m['n_out'] * 4 * m['d_out']
                tot_s += sec
                n += 1
        if tot_s == 0:
            return None
        ach = tot_b / tot_s / 1e9
        return dict(kernel='tp_scatter_kernel (all layers)', bound='hbm', achieved=ach, peak=peak_gbs, unit='GB/s',
                    frac=ach / peak_gbs, traffic=None, peak_source=f'{peak_src} HBM copy bandwidth', launches=n,
                    bytes_per_launch=tot_b / n, ms_per_launch=1e3 * tot_s / n)
