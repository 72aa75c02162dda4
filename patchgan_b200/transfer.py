"""Transfer learning from a checkpoint whose layers only partly match (API of /root/reference/patchgan/transfer.py:4-26:
``module.load_transfer_data(state_dict)``, ``InvalidCheckpointError`` when nothing could be used).

Entries are matched by key AND shape; everything that matches is loaded in one ``load_state_dict(strict=False)`` call,
so the parameters keep their storage (the flat optimizer buffers and captured CUDA graphs of a Trainer stay valid) and
the kernels' packed operand copies are refreshed through the usual version check."""
import torch


class InvalidCheckpointError(Exception):
    pass


class Transferable:
    def load_transfer_data(self, state_dict):
        mine = self.state_dict()
        usable = {key: torch.as_tensor(getattr(value, 'data', value)) for key, value in state_dict.items()
                  if key in mine and tuple(getattr(value, 'shape', ())) == tuple(mine[key].shape)}
        if not usable:
            raise InvalidCheckpointError("Could not load transfer weights")
        self.load_state_dict(usable, strict=False)
        print(f"Loaded {len(usable)} weights out of {len(state_dict)}")
