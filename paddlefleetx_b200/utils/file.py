"""Archive/file helpers (reference ppfleetx/utils/file.py: unzip/untar/parse_csv, path helpers)."""
import csv
import os
import shutil
import tarfile
import zipfile


def unzip(zip_path, out_dir=None, delete=False):
    out_dir = out_dir or os.path.dirname(zip_path)
    with zipfile.ZipFile(zip_path) as z:
        for m in z.namelist():
            if os.path.isabs(m) or ".." in m.split("/"):
                raise ValueError(f"unsafe member path {m}")
        z.extractall(out_dir)
    if delete:
        os.remove(zip_path)
    return out_dir


def untar(tar_path, out_dir=None, delete=False):
    out_dir = out_dir or os.path.dirname(tar_path)
    with tarfile.open(tar_path) as t:
        t.extractall(out_dir, filter="data")
    if delete:
        os.remove(tar_path)
    return out_dir


def parse_csv(path, skip_lines=0, delimiter=" ", quotechar="|", quoting=csv.QUOTE_NONE, check_rows=True):
    with open(path, newline="") as f:
        rows = list(csv.reader(f, delimiter=delimiter, quotechar=quotechar, quoting=quoting))[skip_lines:]
    if check_rows and rows and any(len(r) != len(rows[0]) for r in rows):
        raise ValueError("ragged csv")
    return rows


def ensure_dir(path):
    os.makedirs(path, exist_ok=True)
    return path


def atomic_write(path, data: bytes):
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(data)
        f.flush()
        os.fsync(f.fileno())
    os.replace(tmp, path)


def remove_tree(path):
    shutil.rmtree(path, ignore_errors=True)
