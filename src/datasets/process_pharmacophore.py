"""Pharmacophore (.phore) ingestion and the AncPhore scoring hand-off, host side.

Mirrors /root/reference/src/datasets/process_pharmacophore.py: `parse_phore` (:78-152), `parse_phore_line`
(:751-789), `get_phore_graph` (:634-714), `phore_featurizer` (:717-748), `parse_score_file` (:885-925),
`calc_phore_fitting` (:930-1000).  Pure Python / numpy — no RDKit needed for this half of the preprocessing.
"""
import os
import subprocess
from collections import namedtuple

import numpy as np
import torch

PHORE_TYPES = ['MB', 'HD', 'AR', 'PO', 'HA', 'HY', 'NE', 'CV', 'CR', 'XB', 'EX']
phore_feature_dims = ([len(PHORE_TYPES), 2, 2], 2)
ANCPHORE = os.path.join(os.path.dirname(os.path.abspath(__file__)), '../../programs/AncPhore')

Feature = namedtuple('Feature', 'type alpha weight factor xyz has_norm norm_xyz label anchor_weight')
Phore = namedtuple('Phore', 'id features exclusion_volumes')


def parse_phore_line(record, cvs=False):
    """One tab-separated feature record: type alpha weight factor x y z has_norm nx ny nz label anchor_weight."""
    f = record.split('\t')
    if len(f) != 13:
        raise SyntaxError(f'invalid phore feature record: {record!r}')
    return Feature(f[0] if cvs else f[0][:2], float(f[1]), float(f[2]), float(f[3]), tuple(map(float, f[4:7])),
                   bool(int(f[7])), tuple(map(float, f[8:11])), f[11], float(f[12]))


def parse_phore(phore_file, skip_ex=False, cvs=False):
    """All pharmacophore models of a .phore file (id line, feature records, '$$$$' terminator)."""
    if not os.path.exists(phore_file):
        raise FileNotFoundError(f'The specified pharmacophore file (*.phore) is not found: `{phore_file}`')
    phores, cur_id, feats, exs = [], None, [], []
    with open(phore_file) as fh:
        for line in fh:
            rec = line.strip()
            if not rec:
                break
            if cur_id is None:
                cur_id = rec
            elif rec == '$$$$':
                if feats:
                    phores.append(Phore(cur_id, feats, exs))
                cur_id, feats, exs = None, [], []
            else:
                feat = parse_phore_line(rec, cvs)
                if feat.type != 'EX':
                    feats.append(feat)
                elif not skip_ex:
                    exs.append(feat)
    return phores


def phore_featurizer(points):
    """[type idx, index([True, False], is_EX), index([True, False], has_norm), alpha, weight] per point."""
    rows = []
    for p in points:
        t = PHORE_TYPES.index(p.type) if p.type in PHORE_TYPES else len(PHORE_TYPES) - 1
        rows.append([t, 0 if p.type == 'EX' else 1, 0 if p.has_norm else 1, p.alpha, p.weight])
    return torch.tensor(rows, dtype=torch.float32)


def get_phore_graph(phore, graph, consider_ex=True, neighbor_cutoff=5.0, ex_connected=True):
    """Nodes = features then exclusion spheres; edges: features fully connected among themselves, every exclusion
    sphere to all nodes closer than the cutoff (nearest node if none)."""
    pts = list(phore.features) + (list(phore.exclusion_volumes) if consider_ex else [])
    n_feat, n = len(phore.features), len(pts)
    pos = np.asarray([p.xyz for p in pts], dtype=np.float64)
    nrm = np.asarray([np.subtract(p.norm_xyz, p.xyz) if p.has_norm else (0.0, 0.0, 0.0) for p in pts], dtype=np.float64)
    ln = np.linalg.norm(nrm, axis=1)
    ln[ln == 0] = 1
    nrm = nrm / ln[:, None]
    dist = np.linalg.norm(pos[:, None] - pos[None], axis=-1)
    cutoff = float('inf') if neighbor_cutoff is None else neighbor_cutoff
    src, dst = [], []
    for i in range(n):
        if i < n_feat:
            nb = [j for j in range(n_feat) if j != i]
        else:
            nb = [int(j) for j in np.where(dist[i] < cutoff)[0] if j != i]
            if not ex_connected:
                nb = [j for j in nb if j >= n_feat]
        if not nb:
            nb = [int(np.argsort(dist[i])[1])]
        src += [i] * len(nb)
        dst += nb
    ph = graph['phore']
    ph.x = phore_featurizer(pts)
    ph.pos = torch.from_numpy(pos).float()
    ph.norm = torch.from_numpy(nrm).float()
    graph['phore', 'phore_contact', 'phore'].edge_index = torch.tensor([src, dst], dtype=torch.long)
    return graph


# column of an AncPhore .score line (tab separated, no header) by `fitness` (process_pharmacophore.py:885-925)
SCORE_INDEX = {1: -4, 2: -3, 3: -2, 4: -1, 5: -5, 6: -6}


def parse_score_file(score_file, return_all=False, fitness=1):
    """AncPhore .score file -> one float per pose (column chosen by `fitness`), or the five columns [-6:-1] of every
    pose with `return_all`.  None (after printing the reason) when the file cannot be parsed, like the reference."""
    try:
        lines = open(score_file).readlines()
        if return_all:
            return [[float(x) for x in line.strip().split('\t')[-6:-1]] for line in lines]
        return [float(line.strip().split('\t')[SCORE_INDEX[fitness]]) for line in lines]
    except Exception as e:
        print(f'[E] Failed to parse the score file {score_file}.', e)
        return None


def calc_phore_fitting(ligand_file, phore_file, score_file, dbphore_file, log_file, overwrite=False, return_all=False,
                       exVolume_cutoff=500, overlap_coeff=-1, percent_coeff=-1, anchor_coeff=-1, ancphore_path=ANCPHORE,
                       target_fishing=False, fitness=1, timeout=200):
    """Scores the poses of `ligand_file` against `phore_file` with the external AncPhore binary (a closed black box the
    reference shells out to, process_pharmacophore.py:930-1000; same command line, run from the binary's directory).
    Returns the parsed scores, or None when inputs / the binary are missing or no score file appears."""
    ligand_file, phore_file, score_file, dbphore_file, log_file = (
        os.path.abspath(p) for p in (ligand_file, phore_file, score_file, dbphore_file, log_file))
    name = os.path.basename(ligand_file).split('.')[0]
    ancphore_path = os.path.abspath(ancphore_path)
    ok = True
    for what, p in (('ligand file', ligand_file), ('pharmacophore file', phore_file)):
        if not os.path.exists(p):
            ok = False
            print(f"[E] Failed to calculate the fitting score of ligand `{name}`.\nThe {what} `{p}` doesn't exist.")
    if not os.path.exists(ancphore_path):
        ok = False
        print(f'[E] Invalid path to AncPhore program: `{ancphore_path}`')
    fitness = 5 if target_fishing else fitness
    if ok and (overwrite or not os.path.exists(score_file)):
        cmd = [ancphore_path, '-d', ligand_file, '--refphore', phore_file, '--scores', score_file,
               'usedMultiConformerFile', 'formodel']
        if exVolume_cutoff != 500:
            cmd += ['--exvolume_cutoff', str(exVolume_cutoff)]
        for flag, v in (('--overlap_coeff', overlap_coeff), ('--percent_coeff', percent_coeff), ('--anchor_coeff', anchor_coeff)):
            if v != -1:
                cmd += [flag, str(v)]
        if overwrite and os.path.exists(score_file):
            os.remove(score_file)                                  # a stale file must not be parsed as this run's result
        try:
            with open(log_file, 'w') as lf:
                subprocess.run(cmd, stdout=lf, stderr=subprocess.STDOUT, timeout=timeout, check=False,
                               cwd=os.path.dirname(ancphore_path))
        except (subprocess.TimeoutExpired, OSError) as e:
            print(f'[E] Failed to calculate the fitting score of ligand `{name}`.', e)
    if os.path.exists(score_file):
        return parse_score_file(score_file, return_all=return_all, fitness=fitness)
    print(f'[E] No score file generated for {name} and {os.path.basename(phore_file)}')
    return None
