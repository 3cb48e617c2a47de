{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    400,
    800,
    900
  ],
  "chunk_offsets": [
    0,
    37447,
    74840,
    84492
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": true,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 97,
  "sample_rate": 400.0,
  "sha1_compressed": "656d408c52a186a6aa0da763f9abca47ae62858d",
  "sha1_uncompressed": "e30c4a799b1af682f57de1dd5d8b709e6009800f",
  "shape": [
    900,
    97
  ],
  "version": "1.0"
}