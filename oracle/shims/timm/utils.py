"""timm.utils.NativeScaler is imported (never used) by losses.py:9."""


class NativeScaler:  # pragma: no cover - import-only
    pass
