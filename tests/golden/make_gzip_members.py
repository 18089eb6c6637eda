"""Generate tests/golden/gzip_members.json: multi-member gzip files (RFC 1952 2.2) written by
Python's gzip / zlib modules (zlib 1.3), with what gzip(1) / zlib's gzread make of them, and what
the compiled reference (single member, src/zstrm.c:626-667) makes of the same bytes.

    python tests/golden/make_gzip_members.py

Fixtures are small (a few KB of base64); the script and its output are both committed.
"""
import base64
import gzip
import json
import sys
import zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from support import Corpus  # noqa: E402


def member(data, level=6, name=None):
    import io
    buf = io.BytesIO()
    with gzip.GzipFile(filename=name or "", mode="wb", fileobj=buf, compresslevel=level, mtime=0) as f:
        f.write(data)
    return buf.getvalue()


def main():
    c = Corpus()
    parts = [c.fill(0, 3000, offset=0), c.fill(4, 1800, offset=5000), b"", c.fill(2, 2500, offset=100), b"x"]
    cases = {}
    ms = [member(parts[0], 6), member(parts[1], 9, name="second.json"), member(parts[2], 1), member(parts[3], 1), member(parts[4], 6)]
    cases["five_members"] = {"gz": b"".join(ms), "want": b"".join(parts), "members": 5}
    cases["two_members_then_garbage"] = {"gz": ms[0] + ms[1] + b"\x00\x00trailing bytes that are not a member", "want": parts[0] + parts[1], "members": 2}
    cases["single_member"] = {"gz": ms[0], "want": parts[0], "members": 1}
    cases["second_member_truncated"] = {"gz": ms[0] + ms[1][:-6], "want": None, "members": 1, "prefix": parts[0]}
    cases["second_member_bad_crc"] = {"gz": ms[0] + ms[1][:-8] + bytes([ms[1][-8] ^ 1]) + ms[1][-7:], "want": None, "members": 1, "prefix": parts[0]}
    out = {}
    for k, v in cases.items():
        if k == "two_members_then_garbage":
            # gzip(1) decodes the members and warns "trailing garbage ignored"; Python's reader raises
            assert gzip.decompress(ms[0] + ms[1]) == v["want"]
        elif v["want"] is not None:
            assert gzip.decompress(v["gz"]) == v["want"], k       # what zlib's gzip reader makes of it
        else:
            try:
                gzip.decompress(v["gz"])
                raise SystemExit("expected a failure: " + k)
            except (EOFError, gzip.BadGzipFile, zlib.error):
                pass
        out[k] = {"gz": base64.b64encode(v["gz"]).decode(), "members": v["members"],
                  "crc32": zlib.crc32(v["want"]) if v["want"] is not None else None,
                  "size": len(v["want"]) if v["want"] is not None else None,
                  "first_member_crc32": zlib.crc32(parts[0]), "first_member_size": len(parts[0])}
    (ROOT / "tests" / "golden" / "gzip_members.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
