"""The restatement's reading of the GAF columns (rcls.rs:119-146: tab separated, no header, no quoting, `*` = null in every
column, lines starting with `@` skipped, columns 1, 2, 6, 7, 8, 9, 12) against an independent CSV engine: Apache Arrow's reader
with the same options.  Not the reference's reader (polars 0.46 cannot run here), but code this repo did not write: the column
split, the null rule and the integer conversion of the synthetic vg-giraffe and GraphAligner dialects must agree cell for cell."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.csv as pc
import pytest

from common import NASTY, NASTY_DUP, opy, synth


def arrow_columns(gaf: bytes):
    # polars' comment_prefix drops the lines that START with the prefix; Arrow has no such option, so they are removed here
    body = b"\n".join(l for l in gaf.split(b"\n") if not l.startswith(b"@"))
    ncol = max(l.count(b"\t") + 1 for l in body.split(b"\n") if l)
    names = [f"column_{i + 1}" for i in range(ncol)]
    types = {n: pa.string() for n in names}
    for k in (2, 7, 8, 9, 12):
        types[f"column_{k}"] = pa.int64()
    t = pc.read_csv(io.BytesIO(body), read_options=pc.ReadOptions(column_names=names),
                    parse_options=pc.ParseOptions(delimiter="\t", quote_char=False, escape_char=False, newlines_in_values=False),
                    convert_options=pc.ConvertOptions(column_types=types, null_values=["*"], strings_can_be_null=True, quoted_strings_can_be_null=True))
    return {k: t.column(f"column_{k}").to_pylist() for k in (1, 2, 6, 7, 8, 9, 12)}


@pytest.mark.parametrize("params,n", [(NASTY, 4000), (NASTY_DUP, 4000), (synth.GafParams(long_reads=True, id_pair_suffix=False), 300)])
def test_columns_agree_with_arrow(params, n):
    ds = synth.Dataset(31, [3000, 800], [4, 2])
    gaf = ds.gaf(7, 0, n, params)
    # Arrow wants rectangular input: give every line the same number of columns (the generators write 12 mandatory fields + tags)
    lines = [l for l in gaf.split(b"\n") if l]
    width = max(l.count(b"\t") for l in lines if not l.startswith(b"@"))
    gaf = b"\n".join(l if l.startswith(b"@") else l + b"\tx" * (width - l.count(b"\t")) for l in lines) + b"\n"
    rows = opy.load_gaf(gaf)
    cols = arrow_columns(gaf)
    assert len(rows) == len(cols[1]) > 0.9 * n
    dec = lambda v: None if v is None else v.decode()
    assert [dec(r.read_id) for r in rows] == cols[1]
    assert [r.read_len for r in rows] == cols[2]
    assert [dec(r.path) for r in rows] == cols[6]
    assert [r.read_path_len for r in rows] == cols[7]
    assert [r.read_start for r in rows] == cols[8]
    assert [r.read_end for r in rows] == cols[9]
    assert [r.mapq for r in rows] == cols[12]
    # the hostile cases are really in there
    if params is not synth.GafParams and n >= 4000:
        assert any(v is None for v in cols[6]) and any(v is None for v in cols[9])
