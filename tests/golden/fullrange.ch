{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    128,
    256,
    300
  ],
  "chunk_offsets": [
    0,
    8459,
    16918,
    19833
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 33,
  "sample_rate": 128.0,
  "sha1_compressed": "fc75fa1b8fc3b28949e33ad6a0e76a4b97136f74",
  "sha1_uncompressed": "6d2f2ea0ccd83745b28cd086856dec8d27a26b1c",
  "shape": [
    300,
    33
  ],
  "version": "1.0"
}