"""The two helpers of data/devkits/dota_utils.py the merge step uses (:25-40)."""
import os


def custombasename(fullname):
    """'a/b/Task1_plane.txt' -> 'Task1_plane' (dota_utils.py:25-26)."""
    return os.path.basename(os.path.splitext(fullname)[0])


def GetFileFromThisRootDir(dir, ext=None):
    """Every file under `dir` (recursively); `ext`: keep the extensions contained in it (dota_utils.py:29-40)."""
    out = []
    for root, _, files in os.walk(dir):
        for name in files:
            path = os.path.join(root, name)
            if ext is None or os.path.splitext(path)[1][1:] in ext:
                out.append(path)
    return out
