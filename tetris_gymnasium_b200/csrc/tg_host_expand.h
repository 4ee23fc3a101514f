// tg_host_expand.h -- host side of the compact host-buffer step (tg_step_host, TG_HOST_COMPACT).
//
// The observation dict of Tetris._get_obs (envs/tetris.py:566-615) is a pure function of the env's packed state: the hot
// record (position, piece, rotation, holder, queue) and the nibble-packed id plane of its board record.  Instead of moving
// the 2*Hp*Wp + 16 + 16Q bytes of the dict over PCIe, the compact path moves the packed records (hot 32 B + board record)
// and expands them into the caller's host arrays with streaming (non-temporal) stores on a pool of host threads.  This is
// a FORMAT CONVERSION of device results, not a CPU implementation of the game: no game logic lives here.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace tgh {

struct ExpandCfg {
    int W, H, Wp, Hp, Q;
    int OB, OQ;          // bytes per env of the board / mask image and of the queue image
    int holder_size, OH; // TetrominoHolder(size) and bytes per env of the holder image (16 * size)
    int hdr;             // packed link records: bytes of hot words in front of the id plane (12, or 16 with the holder FIFO word)
    int board_stride;    // bytes per env of the packed board record
    int ids_off;         // byte offset of the nibble id plane inside a board record
    int ids_bytes;       // bytes of the id plane that hold cells (ceil(H * W / 2))
    uint32_t rowbytes[8][4][4];   // [piece][rotation][matrix row] -> four id-valued bytes (piece 7 = zeros: unused nibble)
    uint16_t cells[8][4];         // [piece][rotation] -> four cells, nibble k = (i << 2) | j
    int n[8];                     // matrix size of a piece
    // byte-permute tables of the vector path (AVX-512 VBMI): output vector v = bytes [64v, 64v + 64) of the board image; its
    // cell bytes come from the 64-byte window of the id plane starting at vbase[v]
    int nvec, vec_ok, vreach;     // vectors per image; tables valid; bytes of the id plane the windows read (from ids)
    int vbase[32];
    uint64_t vodd[32], vcell[32]; // per byte: takes the high nibble / is a playfield cell (else bedrock)
    alignas(64) uint8_t vidx[32][64];
};
void build_vector_tables(ExpandCfg& c);   // fills nvec .. vidx from W, H

struct ExpandArgs {
    // state records: either the device layout (hot [n][32] + board [n][board_stride], ids at ids_off) or, with hot == nullptr,
    // the packed link format (board = [n][board_stride] with hot words 0, 2, 3 in the first 12 bytes and the id plane behind:
    // the caller passes an ExpandCfg whose board_stride / ids_off describe that record)
    const uint8_t* hot;
    const uint8_t* board;
    const uint8_t* board_end;   // first byte after the readable board records (the vector path reads whole 64-byte windows)
    uint8_t* o_board;       // [n][Hp][Wp]
    uint8_t* o_mask;        // [n][Hp][Wp]
    uint8_t* o_holder;      // [n][4][4]
    uint8_t* o_queue;       // [n][4][4Q]
};

// expands envs [e0, e1) on the calling thread
void expand_range(const ExpandCfg& c, const ExpandArgs& a, int64_t e0, int64_t e1);

// persistent worker pool (the calling thread takes part): run(n_items, fn) calls fn(item) for item = 0 .. n_items-1
class Pool {
  public:
    explicit Pool(int threads);
    ~Pool();
    int size() const { return nthreads_; }
    void run(int64_t n_items, void (*fn)(void* ctx, int64_t item), void* ctx);

  private:
    struct Impl;
    Impl* impl_;
    int nthreads_;
};

// streaming-store fill of [dst, dst + bytes) (64-byte aligned): the ceiling the expansion's output traffic is measured against
void stream_fill(uint8_t* dst, size_t bytes, int value);

int default_threads();   // cores this process may run on / LOCAL_WORLD_SIZE (torchrun), overridden by TG_HOST_THREADS
const char* isa_name();  // which row-expansion / store variant the dispatcher picked

}  // namespace tgh
