{
  "algorithm": "zlib",
  "chunk_bounds": [
    0,
    1
  ],
  "chunk_offsets": [
    0,
    10
  ],
  "chunk_order": "F",
  "comp_level": -1,
  "do_spatial_diff": false,
  "do_time_diff": true,
  "dtype": "int16",
  "n_channels": 1,
  "sample_rate": 1.0,
  "sha1_compressed": "e23811cf64e6f2613aae33852f157b88d86d1091",
  "sha1_uncompressed": "6acd00890d8cf3d429041fd803b558544886b474",
  "shape": [
    1,
    1
  ],
  "version": "1.0"
}