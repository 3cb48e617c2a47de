{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    1000,
    2000,
    2500
  ],
  "chunk_offsets": [
    0,
    38929,
    77785,
    97707
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": true,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 40,
  "sample_rate": 1000.0,
  "sha1_compressed": "546e6c3449f0eccc07c8c006500968f05e2bcc43",
  "sha1_uncompressed": "3e219b584766b271cc7b4da961d9d4f7a296cddb",
  "shape": [
    2500,
    40
  ],
  "version": "1.0"
}