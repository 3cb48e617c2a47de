{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    500,
    1000,
    1500
  ],
  "chunk_offsets": [
    0,
    32205,
    64400,
    96643
  ],
  "chunk_order": "C",
  "comp_level": -1,
  "do_spatial_diff": true,
  "do_time_diff": true,
  "dtype": "float64",
  "n_channels": 9,
  "sample_rate": 500.0,
  "sha1_compressed": "70762ad19ce7bcdfe278f2f17b0d114e6b0555f4",
  "sha1_uncompressed": "ede8329130edba82fe3c4b183ceaa1c89b579a77",
  "shape": [
    1500,
    9
  ],
  "version": "1.0"
}