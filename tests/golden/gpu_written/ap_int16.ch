{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    1000,
    2000,
    2600
  ],
  "chunk_offsets": [
    0,
    38335,
    77099,
    100829
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 50,
  "sample_rate": 1000.0,
  "sha1_compressed": "02fd8abf584743e623735aa96045c23d79f0899a",
  "sha1_uncompressed": "5aa54da682e5042a4671d3e040a0967d68a38d0f",
  "shape": [
    2600,
    50
  ],
  "version": "1.0"
}