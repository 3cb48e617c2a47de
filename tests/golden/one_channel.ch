{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    2000,
    4000,
    5000
  ],
  "chunk_offsets": [
    0,
    1812,
    3591,
    4528
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 1,
  "sample_rate": 2000.0,
  "sha1_compressed": "b5de668e7d6e165caa478b7a344f25552c314013",
  "sha1_uncompressed": "6e3509a8615746fd40d298597849a5388afb903b",
  "shape": [
    5000,
    1
  ],
  "version": "1.0"
}