"""vlnce_b200 — B200-native (sm_100a) implementation of the VLN-CE policy hot path behind
the reference's ``Policy.act() / build_distribution()`` surface.  See DESIGN.md."""
from .aux_losses import AuxLosses  # noqa: F401


def __getattr__(name):  # lazy: importing the package must not require torchvision / CUDA
    if name == "CMAPolicy":
        from .cma_policy import CMAPolicy

        return CMAPolicy
    if name == "Seq2SeqPolicy":
        from .seq2seq_policy import Seq2SeqPolicy

        return Seq2SeqPolicy
    if name == "WaypointPolicy":
        from .waypoint_policy import WaypointPolicy

        return WaypointPolicy
    raise AttributeError(name)
