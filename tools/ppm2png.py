import sys, zlib, struct
import numpy as np
def read_ppm(path):
    t = open(path).read().split()
    w, h = int(t[1]), int(t[2])
    return np.array(t[4:], dtype=np.uint8).reshape(h, w, 3)
def write_png(path, img):
    h, w, _ = img.shape
    raw = b''.join(b'\x00' + img[y].tobytes() for y in range(h))
    def chunk(tag, data):
        c = struct.pack('>I', len(data)) + tag + data
        return c + struct.pack('>I', zlib.crc32(tag + data) & 0xffffffff)
    open(path, 'wb').write(b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, 2, 0, 0, 0)) + chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))
for p in sys.argv[1:]:
    write_png(p.replace('.ppm', '.png'), read_ppm(p))
