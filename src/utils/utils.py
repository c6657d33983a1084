"""utils.utils.get_model of the reference (/root/reference/src/utils/utils.py:113-168): Namespace -> constructor
kwargs, identical mapping (note: `multiple` is NOT forwarded there either)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from models.score_model_phore import TensorProductScoreModel as PhoreModel   # noqa: E402
from utils.diffusion_utils import get_timestep_embedding                      # noqa: E402


def get_model(args, device, t_to_sigma, no_parallel=False, confidence_mode=False, contrastive_model=None):
    if getattr(args, 'model_type', 'diff') != 'diff':
        raise NotImplementedError("only model_type == 'diff' exists on the B200 path")
    timestep_emb_func = get_timestep_embedding(args.embedding_type, args.sigma_embed_dim, args.embedding_scale)
    model = PhoreModel(t_to_sigma=t_to_sigma, device=device, no_torsion=args.no_torsion,
                       timestep_emb_func=timestep_emb_func, num_conv_layers=args.num_conv_layers,
                       lig_max_radius=args.max_radius, scale_by_sigma=args.scale_by_sigma,
                       sigma_embed_dim=args.sigma_embed_dim, ns=args.ns, nv=args.nv,
                       distance_embed_dim=args.distance_embed_dim,
                       cross_distance_embed_dim=args.cross_distance_embed_dim, batch_norm=not args.no_batch_norm,
                       dropout=args.dropout, use_second_order_repr=args.use_second_order_repr,
                       cross_max_distance=args.cross_max_distance, dynamic_max_cross=args.dynamic_max_cross,
                       confidence_mode=confidence_mode, consider_norm=args.consider_norm,
                       use_phore_rule=args.phore_rule, auto_phorefp=args.auto_phorefp, angle_match=args.angle_match,
                       cross_distance_transition=args.cross_distance_transition,
                       phore_direction_transition=args.phore_direction_transition,
                       phoretype_match_transition=args.phoretype_match_transition, new=args.new,
                       ex_factor=args.ex_factor, boarder=getattr(args, 'boarder', False),
                       by_radius=getattr(args, 'by_radius', False),
                       clash_tolerance=getattr(args, 'clash_tolerance', 0.4),
                       clash_cutoff=getattr(args, 'clash_cutoff', [1.0, 2.0, 3.0, 4.0, 5.0]),
                       use_att=getattr(args, 'use_att', False),
                       use_phore_match_feat=getattr(args, 'use_phore_match_feat', False),
                       atom_weight=getattr(args, 'atom_weight', 'softmax'),
                       trioformer_layer=getattr(args, 'trioformer_layer', 1), contrastive_model=contrastive_model,
                       scaler=getattr(args, 'scaler', 1.0))
    model.to(device)
    return model
