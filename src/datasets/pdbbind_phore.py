"""datasets.pdbbind_phore of the reference (/root/reference/src/datasets/pdbbind_phore.py), hot-path part only:
`NoiseTransformPhore.sample_from_infer` (:286-359), the "calibrated sampler" step of training-time augmentation - one denoising
step of the current model from x_(n+1), re-expressed as an update of the clean pose x_0.  The score model and both conformer
updates run on the GPU (utils.sampling.sample_step / apply_perturbations); everything else of the class (dataset caches, RDKit
preprocessing, apply_noise) is out of scope."""
import copy
import os

import numpy as np
import torch

from diffphore_b200.graph import collate
from utils import so3, torus
from utils.diffusion_utils import set_time_phore
from utils.sampling import sample_step, get_updates_from_0_to_n, apply_perturbations


class NoiseTransformPhore:
    def __init__(self, t_to_sigma, no_torsion, epochs=None, reject=False, cofactor=0.3, calc_fitscore=False,
                 fitscore_tmp='/tmp/diffphore/fitscore_tmp/', delta_t=0.05, rate_from_infer=0.1, epoch_from_infer=300,
                 dynamic_coeff=0, model=None, args=None, **kwargs):
        self.t_to_sigma, self.no_torsion, self.epochs, self.reject = t_to_sigma, no_torsion, epochs, reject
        self.current_epoch, self.cofactor = 0, cofactor
        self.calc_fitscore, self.fitscore_tmp = calc_fitscore, fitscore_tmp
        self.delta_t = self.delta_t0 = delta_t
        self.rate_from_infer, self.epoch_from_infer, self.dynamic_coeff, self.p = rate_from_infer, epoch_from_infer, dynamic_coeff, None
        # the reference keeps a CPU deep copy of the model for this; here the model stays where its kernels are
        self.model = (model if not hasattr(model, 'module') else model.module) if (rate_from_infer > 0 and model is not None) else None
        self.args = args

    def __call__(self, data):
        raise NotImplementedError('training-time noise transform (apply_noise) is out of scope of the B200 hot path (SURVEY 2.1)')

    def get_fitscore(self, data):
        """AncPhore fitness of the current pose (reference :233-284); only with calc_fitscore=True."""
        if not self.calc_fitscore:
            return
        from datasets.process_mols import write_mol_with_multi_coords
        from datasets.process_pharmacophore import calc_phore_fitting
        name = data.name
        lig_pos = data['ligand'].pos.numpy() + data.original_center.numpy()
        keep = torch.not_equal(data['ligand'].x[:, 0], 0).cpu().numpy()
        tmp_out = os.path.join(self.fitscore_tmp, name)
        os.makedirs(tmp_out, exist_ok=True)
        lig_file = os.path.join(tmp_out, 'ligand.sdf')
        write_mol_with_multi_coords(data.sdf_template, lig_pos[keep][None], lig_file, name)
        scores = calc_phore_fitting(lig_file, data.phore_file, os.path.join(tmp_out, f'{name}.score'), os.path.join(tmp_out, f'{name}.dbphore'),
                                    os.path.join(tmp_out, f'{name}.log'), overwrite=True, return_all=True)
        if not scores:
            print(data.name, scores)
        data.fitscore, data.ph_overlap, data.ex_overlap = scores[0]

    def sample_from_infer(self, data_0, data, t_n_1, tr_sigma, rot_sigma, tor_sigma, torsion_updates=None, debug=False):
        """x_(n+1) -> x_n with the current model (sample_step on the GPU), then the update (translation, rotation vector, torsions)
        that carries the clean pose data_0 onto x_n, applied to data_0 (conformer update on the GPU), with the scores of that
        update as training targets (reference :286-359)."""
        result, info = data, {}
        if self.model is not None and self.args is not None and data_0 is not None:
            batch = collate([copy.deepcopy(data)])
            for key in ('node_t',):                          # the times set by the caller travel with the graph
                pass
            if 'complex_t' in data:
                batch.complex_t = {k: torch.as_tensor(v).reshape(-1)[:1] for k, v in data.complex_t.items()}
            _data, tor_p, tr_p, rot_p = sample_step(batch, self.model, self.args, tr_sigma, rot_sigma, tor_sigma, delta_t=self.delta_t)
            _data = _data[0]
            tor_up = np.zeros(int(_data['ligand'].edge_mask.sum()), dtype=np.float64)
            if torsion_updates is not None:
                tor_up += torsion_updates
            if tor_p is not None:
                tor_up += tor_p
            tr_up, rot_up = get_updates_from_0_to_n(data_0, _data, tor_up)
            t = t_n_1 - self.delta_t
            tr_sigma, rot_sigma, tor_sigma = self.t_to_sigma(t, t, t)
            set_time_phore(data_0, t, t, t, 1, 'cpu')
            b0 = collate([data_0])
            moved = apply_perturbations(self.model, b0, tr_up.float(), torch.from_numpy(rot_up).float()[None], tor_up)[0]
            data_0['ligand'].pos, data_0['ligand'].norm = moved['ligand'].pos, moved['ligand'].norm
            self.get_fitscore(data_0)
            data_0.tr_score = -tr_up / tr_sigma ** 2
            data_0.rot_score = torch.from_numpy(so3.score_vec(vec=rot_up, eps=rot_sigma)).float().unsqueeze(0)
            data_0.tor_score = None if self.no_torsion else torch.from_numpy(torus.score(tor_up, tor_sigma)).float()
            data_0.tor_sigma_edge = None if self.no_torsion else np.ones(int(data_0['ligand'].edge_mask.sum())) * tor_sigma
            result = data_0
            if debug:
                info.update(tor_perturb_n_1_n=tor_p, tr_perturb_n_1_n=tr_p, rot_perturb_n_1_n=rot_p, tor_perturb_0_n=tor_up,
                            tr_perturb_0_n=tr_up, rot_perturb_0_n=rot_up,
                            rmsd=((data_0['ligand'].pos - _data['ligand'].pos) ** 2).sum(dim=1).mean().sqrt())
        return result, info

    def from_infer(self, t):
        if self.model is not None and t > self.delta_t:
            if self.dynamic_coeff == 0:
                return self.current_epoch >= self.epoch_from_infer and np.random.uniform() < self.rate_from_infer
            return np.random.uniform() < self.p
        return False

    def update_model(self, state_dict):
        if self.model is not None:
            self.model.load_state_dict(state_dict)

    def dynamic_schedule(self, epoch, max_rate=0.4, u=400, c=10):
        return max_rate * (1 - u / (u + np.exp(c * epoch / u)))
