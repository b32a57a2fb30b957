from .checkpoint import load_policy_from_checkpoint, save_policy_checkpoint

__all__ = ["load_policy_from_checkpoint", "save_policy_checkpoint"]
