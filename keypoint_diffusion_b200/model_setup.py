"""model_from_config: YAML config dict -> KeypointDiffusion, with the reference's mapping of
config sections to constructor kwargs (reference model_setup.py:4-64)."""
import copy
from pathlib import Path

import torch
import yaml

from .ligand_diffuser import KeypointDiffusion


def model_from_config(config: dict) -> KeypointDiffusion:
    config = copy.deepcopy(config)
    architecture = config['diffusion'].get('architecture', 'egnn')
    rec_encoder_type = config['diffusion'].get('rec_encoder_type', 'learned')
    use_fake_atoms = config['dataset'].get('max_fake_atom_frac', 0) > 0
    n_rec_feat = len(config['dataset']['rec_elements'])
    n_lig_feat = len(config['dataset']['lig_elements']) + (1 if use_fake_atoms else 0)
    if rec_encoder_type == 'learned':
        n_kp_feat = (config['rec_encoder']['out_n_node_feat'] if architecture == 'egnn'
                     else config['rec_encoder_gvp']['out_scalar_size'])
    else:
        n_kp_feat = n_rec_feat
    if architecture == 'gvp':
        rec_encoder_config = config['rec_encoder_gvp']
        rec_encoder_config['in_scalar_size'] = n_rec_feat
        dynamics_config = config['dynamics_gvp']
    else:
        rec_encoder_config = config['rec_encoder']
        rec_encoder_config['in_n_node_feat'] = n_rec_feat
        dynamics_config = config['dynamics']
    return KeypointDiffusion(n_lig_feat, n_kp_feat, processed_dataset_dir=Path(config['dataset']['location']),
                             graph_config=config['graph'], dynamics_config=dynamics_config,
                             rec_encoder_config=rec_encoder_config,
                             rec_encoder_loss_config=config.get('rec_encoder_loss', {}),
                             use_fake_atoms=use_fake_atoms, **config['diffusion'])


def load_model(model_dir, device="cuda", checkpoint: str = "model.pt") -> KeypointDiffusion:
    """model_dir/config.yml + model_dir/model.pt (a plain state_dict), as the reference's test.py:91-127.
    With a checkpoint present, the keypoint feature width follows the checkpoint."""
    model_dir = Path(model_dir)
    with open(model_dir / 'config.yml') as f:
        config = yaml.safe_load(f)
    ckpt = model_dir / checkpoint
    sd = torch.load(ckpt, map_location='cpu') if ckpt.exists() else None
    if sd is not None and config['diffusion'].get('rec_encoder_type', 'learned') == 'fixed':
        for key in ('dynamics.rec_encoder.0.weight', 'dynamics.kp_encoder.0.weight'):
            if key in sd:
                width = sd[key].shape[1] - (1 if 'kp_encoder' in key else 0)
                config['dataset']['rec_elements'] = list(range(width))
    model = model_from_config(config)
    if sd is not None:
        model.load_state_dict(sd, strict=True)
    return model.to(device).eval()
