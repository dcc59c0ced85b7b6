"""The two helpers of the reference's utilities/misc.py that test_diml_cvt.py uses
(:16,:89,:135): parameter count and checkpoint loading."""
import numpy as np
import torch


def gimme_params(model):
    """utilities/misc.py:8-11: number of trainable weights."""
    return sum(int(np.prod(p.size())) for p in model.parameters() if p.requires_grad)


def load_checkpoint(model, optimizer, save_path):
    """utilities/misc.py:54-69: loads {'model', 'optimizer', 'best_metrics', 'epoch'}; strips the
    DataParallel 'module.' prefix; returns (best_metrics, epoch)."""
    print('Load checkpoint from', save_path)
    state = torch.load(save_path, map_location='cpu')
    weights = {(k[7:] if k.startswith('module') else k): v for k, v in state['model'].items()}
    model.load_state_dict(weights)
    if optimizer is not None:
        optimizer.load_state_dict(state['optimizer'])
    return state['best_metrics'], state.get('epoch', 0)
