"""Cached download helper (reference ppfleetx/utils/download.py: cached_path/get_path_from_url with md5 check and
rank-0-downloads-others-wait).  Network may be absent: local paths and ``file://`` URLs always work, remote URLs raise a clear
error after the cache miss instead of hanging."""
import hashlib
import os
import shutil
import time
import urllib.parse
import urllib.request

from .log import logger

CACHE_HOME = os.path.expanduser(os.environ.get("PFX_HOME", "~/.cache/paddlefleetx_b200"))


def is_url(path):
    return isinstance(path, str) and path.startswith(("http://", "https://", "file://"))


def _md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def _download(url, dst, timeout=30):
    tmp = dst + ".part"
    parsed = urllib.parse.urlparse(url)
    if parsed.scheme == "file":
        shutil.copyfile(parsed.path, tmp)
    else:
        with urllib.request.urlopen(url, timeout=timeout) as r, open(tmp, "wb") as f:
            shutil.copyfileobj(r, f)
    os.replace(tmp, dst)


def get_path_from_url(url, root_dir=None, md5sum=None, check_exist=True):
    root_dir = root_dir or CACHE_HOME
    os.makedirs(root_dir, exist_ok=True)
    dst = os.path.join(root_dir, os.path.basename(urllib.parse.urlparse(url).path))
    rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", 0)))
    if check_exist and os.path.exists(dst) and (md5sum is None or _md5(dst) == md5sum):
        return dst
    if rank == 0:
        logger.info(f"downloading {url} -> {dst}")
        try:
            _download(url, dst)
        except Exception as exc:
            raise RuntimeError(f"cannot fetch {url}: {exc}. Place the file at {dst} manually (offline environment).") from exc
        if md5sum is not None and _md5(dst) != md5sum:
            raise RuntimeError(f"md5 mismatch for {dst}")
    else:
        t0 = time.time()
        while not os.path.exists(dst):
            if time.time() - t0 > 3600:
                raise TimeoutError(f"waited 1h for rank 0 to download {url}")
            time.sleep(1)
    return dst


def cached_path(url_or_path, cache_dir=None, md5sum=None):
    if is_url(url_or_path):
        return get_path_from_url(url_or_path, cache_dir, md5sum)
    if os.path.exists(url_or_path):
        return url_or_path
    raise FileNotFoundError(url_or_path)


def download(url, path):
    """Fetch ``url`` to the exact file ``path``: local rank 0 downloads, the other ranks of the node wait for the file to appear
    (reference utils/download.py:117-128)."""
    rank, world = int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and rank != 0:
        t0 = time.time()
        while not os.path.exists(path):
            if time.time() - t0 > 3600:
                raise TimeoutError(f"waited 1h for local rank 0 to download {url}")
            time.sleep(1)
        return path
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    last = None
    for _ in range(3):                      # the reference retries three times
        try:
            _download(url, path)
            return path
        except Exception as exc:            # noqa: BLE001 - reported below with the target path
            last = exc
    raise RuntimeError(f"cannot fetch {url}: {last}. Place the file at {path} manually (offline environment).") from last
