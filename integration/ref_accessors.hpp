// integration/ref_accessors.hpp -- the reference-side diff the C++ binding needs, stated as code.
//
// The reference keeps the pools, the geometry and the page-table dimensions PRIVATE and offers no accessors
// (kv_cache/kv_tile_cache.hpp:43-51, kv_cache/page_table.hpp:24-33), so a launcher outside those classes cannot name
// the buffers it has to hand to the kernels.  A maintainer adds the eight one-line const accessors below
// (marked ADD) -- nothing else in the two headers changes.  For the compile check in tests/test_integration_stub.py this file
// MIRRORS the declarations of the two reference classes member for member (same names, same order, same types;
// method bodies omitted because they do not compile as shipped) and adds the accessors.
#pragma once
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

class PageTable {  // kv_cache/page_table.hpp:5-37
public:
    PageTable();
    ~PageTable();
    void init(int num_beams, int num_heads, int num_tiles);
    void clear();
    void assign(int beam_id, int head_id, int tile_id, int page_id);
    int* device_data() const;
    void sync_to_gpu();
    void remove(int beam_id, int head_id, int tile_id);
    // ADD (page_table.hpp, public): dimensions of the dense table
    int num_beams() const { return num_beams_; }
    int num_heads() const { return num_heads_; }
    int num_tiles() const { return num_tiles_; }

private:
    int num_beams_;
    int num_heads_;
    int num_tiles_;
    int total_entries_;
    int* d_table_;
    std::vector<int> host_table_;
};

template <typename T>
class KVTileCache {  // kv_cache/kv_tile_cache.hpp:9-80
public:
    KVTileCache();
    ~KVTileCache();
    void init(int num_pages, int tile_size, int head_dim);
    void resize(int new_num_pages, int new_tile_size);
    T* get_key_ptr(int page_id);
    T* get_value_ptr(int page_id);
    void register_tile(int beam_id, int head_id, int tile_id);
    void save_to_file(const std::string& path);
    void load_from_file(const std::string& path);
    void sync_page_table_to_gpu();
    // ADD (kv_tile_cache.hpp, public): what a launcher outside the class must be able to name
    const T* key_buffer() const { return key_buffer_; }
    const T* value_buffer() const { return value_buffer_; }
    int tile_size() const { return tile_size_; }
    int head_dim() const { return head_dim_; }
    int total_pages() const { return total_pages_; }
    const PageTable& page_table() const { return page_table_; }

private:
    T* key_buffer_;
    T* value_buffer_;
    int tile_size_;
    int head_dim_;
    int total_pages_;
    PageTable page_table_;
};
