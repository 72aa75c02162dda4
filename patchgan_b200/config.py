"""YAML configuration shared by patchgan_train / patchgan_infer.  Both schemas found in the reference are accepted:
the nested one read by train.py (train.py:36-127) and the older flat one read by infer.py and used by
examples/train_coco.yaml (infer.py:127-162)."""
import importlib.machinery


def model_params(config):
    """-> dict(gen_filts, activation, use_dropout, final_activation, disc_filts, disc_norm, n_disc_layers)"""
    mp = config['model_params']
    if 'generator' in mp:                                   # nested schema (train.py:85-100)
        g, d = mp['generator'], mp['discriminator']
        return dict(gen_filts=g['filters'], activation=g['activation'], use_dropout=g.get('use_dropout', True),
                    final_activation=g.get('final_activation', 'sigmoid'), disc_filts=d['filters'],
                    disc_norm=d.get('norm', False), n_disc_layers=d['n_layers'])
    return dict(gen_filts=mp['gen_filts'], activation=mp['activation'], use_dropout=mp.get('use_dropout', True),
                final_activation=mp.get('final_activation', 'sigmoid'), disc_filts=mp['disc_filts'],
                disc_norm=mp.get('disc_norm', False), n_disc_layers=mp['n_disc_layers'])


def dataset_class(dataset_params):
    """-> (Dataset class, in_channels, out_channels, extra kwargs)   (train.py:52-68, infer.py:97-117)"""
    if dataset_params['type'] == 'COCOStuff':
        from .io import COCOStuffDataset
        labels = dataset_params.get('labels', [1])
        return COCOStuffDataset, 3, len(labels), {'labels': labels}
    try:
        module = importlib.machinery.SourceFileLoader('io', 'io.py').load_module()
        cls = getattr(module, dataset_params['type'])
    except FileNotFoundError:
        print("Make sure io.py is in the working directory!")
        raise
    except (ImportError, ModuleNotFoundError, AttributeError):
        print(f"io.py does not contain {dataset_params['type']}")
        raise
    return cls, dataset_params.get('in_channels', 3), dataset_params.get('out_channels', 1), {}


def data_paths(config):
    """-> (train paths, validation paths, train_val_split) accepting `dataset.train_data` (nested) or top-level
    `train_data` (flat, examples/train_coco.yaml:5-12)."""
    ds = config['dataset']
    if 'train_data' in ds and 'validation_data' in ds:
        return ds['train_data'], ds['validation_data'], None
    if 'data' in ds and 'train_val_split' in ds:
        return ds['data'], None, ds['train_val_split']
    if 'train_data' in config and 'validation_data' in config:
        return config['train_data'], config['validation_data'], None
    raise AttributeError("Please provide either the training and validation data paths or a train/val split!")
