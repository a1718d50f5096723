// Host-only helper: Castagnoli CRC-32C (slicing-by-8) for the TF checkpoint bundle reader
// (tensor_bundle stores a masked crc32c per tensor and per table block).
#include "../../include/ukbb_fcn.h"

static uint32_t g_tab[8][256];
static bool g_init = false;

static void init_tab() {
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
        g_tab[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
        for (int s = 1; s < 8; ++s) g_tab[s][i] = (g_tab[s - 1][i] >> 8) ^ g_tab[0][g_tab[s - 1][i] & 0xFF];
    g_init = true;
}

extern "C" uint32_t ukbb_crc32c(const void* data, size_t n) {
    if (!g_init) init_tab();
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint32_t c = 0xFFFFFFFFu;
    while (n >= 8) {
        const uint32_t lo = c ^ (p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24);
        const uint32_t hi = p[4] | (uint32_t)p[5] << 8 | (uint32_t)p[6] << 16 | (uint32_t)p[7] << 24;
        c = g_tab[7][lo & 0xFF] ^ g_tab[6][(lo >> 8) & 0xFF] ^ g_tab[5][(lo >> 16) & 0xFF] ^ g_tab[4][lo >> 24] ^
            g_tab[3][hi & 0xFF] ^ g_tab[2][(hi >> 8) & 0xFF] ^ g_tab[1][(hi >> 16) & 0xFF] ^ g_tab[0][hi >> 24];
        p += 8; n -= 8;
    }
    while (n--) c = g_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}
