"""torchsnapshot_b200 — B200-native data plane behind TorchSnapshot's take/async_take/restore API."""
__version__ = "0.1.0"
