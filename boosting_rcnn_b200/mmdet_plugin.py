"""``custom_imports`` target for a real mmdet install (level 1):

    custom_imports = dict(imports=['boosting_rcnn_b200.mmdet_plugin'])

Re-registers the hot-path classes under the reference's registry names
(mmdet/models/builder.py:7-15; plugin loading tools/train.py:94-96).  Set
BRCNN_PLUGIN_CLASSES=ProbRoIHead,SingleRoIExtractor,... to replace a subset.
"""
import os

from .registry import HOT_PATH_CLASSES, register_into_mmdet

_names = os.environ.get('BRCNN_PLUGIN_CLASSES')
REGISTERED = register_into_mmdet(
    force=True, names=tuple(_names.split(',')) if _names else HOT_PATH_CLASSES)
