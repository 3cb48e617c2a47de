{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    150,
    300,
    400
  ],
  "chunk_offsets": [
    0,
    5933,
    11863,
    15901
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": true,
  "do_time_diff": false,
  "dtype": "int16",
  "n_channels": 40,
  "sample_rate": 150.0,
  "sha1_compressed": "6896468ed20be732199919ca4dcd36a5917b04ac",
  "sha1_uncompressed": "db0968cd716a1fa5f0240d0a1a148f29fdf10182",
  "shape": [
    400,
    40
  ],
  "version": "1.0"
}