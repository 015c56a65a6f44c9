// gl_hdf5.h - the part of the HDF5 file format a GADGET snapshot needs, written from the format specification.
//
// The reference writes and reads format 3 through libhdf5 (io.c:2533-2866, read_ic.c:648-1060): one file, a group
// "/Header" whose attributes are the fields of struct io_header, one group "/PartType<t>" per particle type present and
// in it one contiguous dataset per record.  This image has no HDF5 library, so the container itself is produced and
// parsed here - the structures libhdf5 emits for exactly those calls (H5Fcreate / H5Gcreate / H5Acreate / H5Dcreate with
// default property lists):
//   superblock version 0; old-style groups (object header with a symbol-table message -> version-1 B-tree node ->
//   symbol-table node, names in a local heap); version-1 object headers; dataspace version 1, datatype version 1 (fixed
//   point / IEEE float, little endian), fill-value version 1, layout version 3 contiguous, attribute version 1.
// The writer lays the metadata out in front and the raw data behind it; all offsets are known before a payload byte is
// written (the snapshot writer fills the payloads in parallel at their offsets).
// The reader accepts what libhdf5 writes for such files with default settings, old or new enough: superblock 0 / 1,
// B-trees of any depth, object-header continuation blocks, layout versions 1-3 (contiguous or compact), dataspace 1 / 2,
// attribute 1-3, integer and float types of 1-8 bytes in little-endian order; chunked or filtered datasets, new-style
// groups (version-2 object headers) and big-endian data are reported as unsupported, not guessed.
// Checked against a file the HDF5 library itself wrote (tests/test_hdf5.py) and by an independent reader
// (tests/h5mini.py).  Host code only.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <map>

namespace glh5
{
typedef unsigned long long u64;
const u64 UNDEF = ~0ull;
enum Type { T_U32, T_I32, T_U64, T_F32, T_F64 };
inline int type_size(Type t) { return (t == T_U64 || t == T_F64) ? 8 : 4; }

// ------------------------------------------------------------------------------------------------- writer ----
struct Attr
{
  std::string name;
  Type type;
  int count;			// 0 = scalar dataspace, otherwise a rank-1 dataspace of that many elements
  std::vector<char> data;
};
struct Dataset
{
  std::string name;
  Type type;
  int rank;			// 1 or 2
  u64 dims[2];
  u64 data_off = 0, bytes = 0;	// out: where the raw data go
};
struct Group
{
  std::string name;
  std::vector<Attr> attrs;
  std::vector<Dataset> dsets;
};

class Writer
{
  std::vector<char> &m;
  static u64 pad8(u64 x) { return (x + 7) & ~7ull; }
  void put(u64 at, const void *p, size_t n)
  {
    if(m.size() < at + n) m.resize(at + n, 0);
    memcpy(m.data() + at, p, n);
  }
  template <class T> void put_v(u64 at, T v) { put(at, &v, sizeof(T)); }
  u64 alloc(u64 n)
  {
    const u64 at = pad8(m.size());
    m.resize(at + n, 0);
    return at;
  }
  static const int LEAF_K = 16, INT_K = 16;	// symbol-table nodes of 32 entries: one node per group is enough
  static void datatype(std::vector<char> &o, Type t)
  {
    const int sz = type_size(t);
    unsigned char h[8] = { 0, 0, 0, 0, (unsigned char) sz, 0, 0, 0 };
    if(t == T_F32 || t == T_F64)
      {
	h[0] = 0x11;		// version 1, class 1 (floating point)
	h[1] = 0x20;		// little endian, mantissa normalisation: msb implied
	h[2] = (unsigned char) (8 * sz - 1);	// sign bit
	o.insert(o.end(), h, h + 8);
	const unsigned short off = 0, prec = (unsigned short) (8 * sz);
	const unsigned char el = (t == T_F32) ? 23 : 52, es = (t == T_F32) ? 8 : 11, ml = 0, ms = el;
	const unsigned int bias = (t == T_F32) ? 127 : 1023;
	char p[12];
	memcpy(p, &off, 2); memcpy(p + 2, &prec, 2);
	p[4] = (char) el; p[5] = (char) es; p[6] = (char) ml; p[7] = (char) ms;
	memcpy(p + 8, &bias, 4);
	o.insert(o.end(), p, p + 12);
      }
    else
      {
	h[0] = 0x10;		// version 1, class 0 (fixed point)
	h[1] = (t == T_I32) ? 0x08 : 0x00;	// little endian, signed bit
	o.insert(o.end(), h, h + 8);
	const unsigned short off = 0, prec = (unsigned short) (8 * sz);
	char p[4];
	memcpy(p, &off, 2); memcpy(p + 2, &prec, 2);
	o.insert(o.end(), p, p + 4);
      }
  }
  static void dataspace(std::vector<char> &o, int rank, const u64 *dims)
  {
    const unsigned char h[8] = { 1, (unsigned char) rank, 0, 0, 0, 0, 0, 0 };	// version 1, no maximum dimensions
    o.insert(o.end(), h, h + 8);
    for(int k = 0; k < rank; k++) o.insert(o.end(), (const char *) &dims[k], (const char *) &dims[k] + 8);
  }
  struct Msg
  {
    unsigned short type;
    unsigned char flags;
    std::vector<char> data;
  };
  static void padto8(std::vector<char> &v) { v.resize((v.size() + 7) & ~(size_t) 7, 0); }
  // version-1 object header: 16-byte prefix, then the messages (8-byte header + data padded to 8 each)
  u64 object_header(const std::vector<Msg> &msgs)
  {
    u64 size = 0;
    for(const Msg &q : msgs) size += 8 + pad8(q.data.size());
    const u64 at = alloc(16 + size);
    m[at] = 1;
    put_v<unsigned short>(at + 2, (unsigned short) msgs.size());
    put_v<unsigned int>(at + 4, 1);	// object reference count
    put_v<unsigned int>(at + 8, (unsigned int) size);
    u64 p = at + 16;
    for(const Msg &q : msgs)
      {
	put_v<unsigned short>(p, q.type);
	put_v<unsigned short>(p + 2, (unsigned short) pad8(q.data.size()));
	m[p + 4] = (char) q.flags;
	if(!q.data.empty()) put(p + 8, q.data.data(), q.data.size());
	p += 8 + pad8(q.data.size());
      }
    return at;
  }
  // group = symbol-table message -> B-tree node -> one symbol-table node; names in a local heap.  entries sorted by name.
  struct Entry
  {
    std::string name;
    u64 obj;
  };
  void group_tables(std::vector<Entry> entries, u64 &btree, u64 &heap)
  {
    std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.name < b.name; });
    // local heap: offset 0 holds the empty name; every name padded to 8; the rest is one free block
    std::vector<u64> noff;
    u64 used = 8;
    for(const Entry &e : entries)
      {
	noff.push_back(used);
	used += pad8(e.name.size() + 1);
      }
    const u64 seg = pad8(used + 32 < 88 ? 88 : used + 32);
    heap = alloc(32);
    const u64 data = alloc(seg);
    put(heap, "HEAP", 4);
    put_v<u64>(heap + 8, seg);
    put_v<u64>(heap + 16, used);	// head of the free list
    put_v<u64>(heap + 24, data);
    for(size_t k = 0; k < entries.size(); k++) put(data + noff[k], entries[k].name.c_str(), entries[k].name.size() + 1);
    put_v<u64>(data + used, 1);	// free block: no next block (H5HL_FREE_NULL), its size
    put_v<u64>(data + used + 8, seg - used);
    // symbol-table node
    u64 snod = 0;
    if(!entries.empty())
      {
	snod = alloc(8 + 2 * LEAF_K * 40);
	put(snod, "SNOD", 4);
	m[snod + 4] = 1;
	put_v<unsigned short>(snod + 6, (unsigned short) entries.size());
	for(size_t k = 0; k < entries.size(); k++)
	  {
	    const u64 e = snod + 8 + 40 * k;
	    put_v<u64>(e, noff[k]);
	    put_v<u64>(e + 8, entries[k].obj);
	  }
      }
    // B-tree node, type 0 (group), level 0
    btree = alloc(24 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8);
    put(btree, "TREE", 4);
    put_v<unsigned short>(btree + 6, entries.empty() ? 0 : 1);
    put_v<u64>(btree + 8, UNDEF);
    put_v<u64>(btree + 16, UNDEF);
    put_v<u64>(btree + 24, 0);	// key 0: the empty name
    if(!entries.empty())
      {
	put_v<u64>(btree + 32, snod);
	put_v<u64>(btree + 40, noff.back());	// key 1: the largest name in the child
      }
  }
  static Msg symtab_msg(u64 btree, u64 heap)
  {
    Msg q{ 0x0011, 0, std::vector<char>(16) };
    memcpy(q.data.data(), &btree, 8);
    memcpy(q.data.data() + 8, &heap, 8);
    return q;
  }
  static Msg attr_msg(const Attr &a)
  {
    std::vector<char> dt, ds;
    datatype(dt, a.type);
    const u64 dim = (u64) a.count;
    dataspace(ds, a.count ? 1 : 0, &dim);
    Msg q{ 0x000C, 0, {} };
    std::vector<char> &o = q.data;
    const unsigned short nsz = (unsigned short) (a.name.size() + 1), tsz = (unsigned short) dt.size(), ssz = (unsigned short) ds.size();
    const unsigned char h[2] = { 1, 0 };
    o.insert(o.end(), h, h + 2);
    o.insert(o.end(), (const char *) &nsz, (const char *) &nsz + 2);
    o.insert(o.end(), (const char *) &tsz, (const char *) &tsz + 2);
    o.insert(o.end(), (const char *) &ssz, (const char *) &ssz + 2);
    o.insert(o.end(), a.name.c_str(), a.name.c_str() + nsz);
    padto8(o);
    o.insert(o.end(), dt.begin(), dt.end());
    padto8(o);
    o.insert(o.end(), ds.begin(), ds.end());
    padto8(o);
    o.insert(o.end(), a.data.begin(), a.data.end());
    return q;
  }

public:
  explicit Writer(std::vector<char> &meta) : m(meta) {}
  // meta <- every byte of the file in front of the raw data; groups[].dsets[].data_off / bytes <- where the raw data go;
  // returns the length of the file (0: more members in a group than one symbol-table node holds)
  u64 layout(std::vector<Group> &groups)
  {
    // one symbol-table node per group (2 * LEAF_K entries); a snapshot has at most 7 groups and 22 datasets per group
    if(groups.size() > 2 * (size_t) LEAF_K) return 0;
    for(const Group &g : groups)
      if(g.dsets.size() > 2 * (size_t) LEAF_K) return 0;
    m.assign(96, 0);		// superblock, filled in at the end
    // raw data sizes first: the dataset headers hold the addresses, and those lie behind all metadata - whose size does
    // not depend on the addresses.  So: build everything with addresses relative to the start of the data area, then
    // add its offset.
    u64 rel = 0;
    for(Group &g : groups)
      for(Dataset &d : g.dsets)
	{
	  d.bytes = d.dims[0] * (d.rank == 2 ? d.dims[1] : 1) * (u64) type_size(d.type);
	  d.data_off = rel;
	  rel = pad8(rel + d.bytes);
	}
    std::vector<std::pair<u64, u64> > fix;	// (position of an address field, relative address)
    std::vector<Entry> root;
    for(Group &g : groups)
      {
	std::vector<Entry> members;
	for(Dataset &d : g.dsets)
	  {
	    std::vector<Msg> msgs;
	    Msg sp{ 0x0001, 0, {} }, ty{ 0x0003, 1, {} }, fv{ 0x0005, 1, {} }, la{ 0x0008, 0, std::vector<char>(18) };
	    dataspace(sp.data, d.rank, d.dims);
	    datatype(ty.data, d.type);
	    const unsigned char fill[8] = { 1, 2, 2, 1, 0, 0, 0, 0 };	// version 1: allocate late, write if set, defined, size 0
	    fv.data.assign(fill, fill + 8);
	    la.data[0] = 3;	// version 3
	    la.data[1] = 1;	// contiguous
	    memcpy(la.data.data() + 10, &d.bytes, 8);
	    msgs.push_back(fv); msgs.push_back(ty); msgs.push_back(sp); msgs.push_back(la);
	    const u64 oh = object_header(msgs);
	    // the layout message is the last one: its data start 8 bytes behind its message header
	    u64 p = oh + 16;
	    for(size_t k = 0; k + 1 < msgs.size(); k++) p += 8 + pad8(msgs[k].data.size());
	    fix.push_back({ p + 8 + 2, d.data_off });
	    members.push_back({ d.name, oh });
	  }
	u64 bt, hp;
	group_tables(members, bt, hp);
	std::vector<Msg> msgs;
	msgs.push_back(symtab_msg(bt, hp));
	for(const Attr &a : g.attrs) msgs.push_back(attr_msg(a));
	root.push_back({ g.name, object_header(msgs) });
      }
    u64 rbt, rhp;
    group_tables(root, rbt, rhp);
    std::vector<Msg> rmsgs;
    rmsgs.push_back(symtab_msg(rbt, rhp));
    const u64 roh = object_header(rmsgs);
    const u64 data0 = pad8(m.size());
    m.resize(data0, 0);
    for(auto &f : fix) put_v<u64>(f.first, data0 + f.second);
    for(Group &g : groups)
      for(Dataset &d : g.dsets) d.data_off += data0;
    const u64 total = data0 + rel;
    // superblock, version 0
    const unsigned char sig[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };
    put(0, sig, 8);
    m[13] = 8;			// size of offsets
    m[14] = 8;			// size of lengths
    put_v<unsigned short>(16, (unsigned short) LEAF_K);
    put_v<unsigned short>(18, (unsigned short) INT_K);
    put_v<u64>(24, 0);		// base address
    put_v<u64>(32, UNDEF);	// free-space info
    put_v<u64>(40, total);	// end of file
    put_v<u64>(48, UNDEF);	// driver info
    put_v<u64>(56, 0);		// root entry: link name offset
    put_v<u64>(64, roh);
    put_v<unsigned int>(72, 1);	// cached: B-tree and heap of the root group
    put_v<u64>(80, rbt);
    put_v<u64>(88, rhp);
    return total;
  }
};

// ------------------------------------------------------------------------------------------------- reader ----
struct Error
{
  std::string what;
};
struct Message
{
  unsigned short type;
  const unsigned char *data;
  size_t size;
};
struct Value			// an attribute or dataset: element type, shape, where the elements are
{
  int cls = -1;			// 0 integer, 1 float
  int size = 0;
  bool sign = false;
  int rank = 0;
  u64 dims[4] = { 0, 0, 0, 0 };
  const unsigned char *data = nullptr;	// little-endian elements, row-major
  u64 count() const
  {
    u64 n = 1;
    for(int k = 0; k < rank; k++) n *= dims[k];
    return n;
  }
  double as_double(u64 i) const
  {
    const unsigned char *p = data + i * (u64) size;
    if(cls == 1)
      {
	if(size == 4) { float f; memcpy(&f, p, 4); return f; }
	double d; memcpy(&d, p, 8); return d;
      }
    unsigned long long u = 0;
    memcpy(&u, p, (size_t) size);
    if(sign && size < 8 && (u >> (8 * size - 1))) u |= ~0ull << (8 * size);
    return sign ? (double) (long long) u : (double) u;
  }
  unsigned long long as_u64(u64 i) const
  {
    if(cls == 1) return (unsigned long long) as_double(i);
    unsigned long long u = 0;
    memcpy(&u, data + i * (u64) size, (size_t) size);
    return u;
  }
};

class Reader
{
  const unsigned char *b;
  u64 n, base = 0;
  u64 root_oh = 0;
  const unsigned char *at(u64 addr, u64 size) const
  {
    if(addr == UNDEF || base + addr + size > n || base + addr + size < base + addr) throw Error{ "HDF5: address outside the file" };
    return b + base + addr;
  }
  template <class T> static T get(const unsigned char *p)
  {
    T v;
    memcpy(&v, p, sizeof(T));
    return v;
  }
  std::vector<Message> header(u64 addr) const
  {
    const unsigned char *h = at(addr, 16);
    if(memcmp(h, "OHDR", 4) == 0) throw Error{ "HDF5: version-2 object headers (new-style groups, libver='latest') are not supported" };
    if(h[0] != 1) throw Error{ "HDF5: unknown object header version" };
    const int nmsg = get<unsigned short>(h + 2);
    std::vector<std::pair<u64, u64> > blocks;
    blocks.push_back({ addr + 16, get<unsigned int>(h + 8) });
    std::vector<Message> out;
    for(size_t k = 0; k < blocks.size() && (int) out.size() < nmsg; k++)
      {
	const unsigned char *q = at(blocks[k].first, blocks[k].second);
	u64 p = 0;
	while(p + 8 <= blocks[k].second && (int) out.size() < nmsg)
	  {
	    const unsigned short t = get<unsigned short>(q + p), ms = get<unsigned short>(q + p + 2);
	    if(p + 8 + ms > blocks[k].second) throw Error{ "HDF5: object header message runs past its block" };
	    if(t == 0x0010 && ms >= 16) blocks.push_back({ get<u64>(q + p + 8), get<u64>(q + p + 16) });
	    out.push_back({ t, q + p + 8, ms });
	    p += 8 + ms;
	  }
      }
    return out;
  }
  static void parse_type(const unsigned char *d, size_t sz, Value &v)
  {
    if(sz < 8) throw Error{ "HDF5: datatype message too short" };
    const int cls = d[0] & 15;
    if(cls != 0 && cls != 1) throw Error{ "HDF5: only integer and floating-point data are supported" };
    if(d[1] & 1) throw Error{ "HDF5: big-endian data are not supported" };
    v.cls = cls;
    v.size = (int) get<unsigned int>(d + 4);
    v.sign = cls == 0 && (d[1] & 8);
    if(cls == 1 && v.size != 4 && v.size != 8) throw Error{ "HDF5: floating-point type is neither 4 nor 8 bytes" };
    if(cls == 0 && (v.size < 1 || v.size > 8)) throw Error{ "HDF5: integer type wider than 8 bytes" };
  }
  static void parse_space(const unsigned char *d, size_t sz, Value &v)
  {
    if(sz < 4) throw Error{ "HDF5: dataspace message too short" };
    const int ver = d[0], rank = d[1];
    size_t p;
    if(ver == 1) p = 8;
    else if(ver == 2) p = 4;
    else throw Error{ "HDF5: unknown dataspace version" };
    if(rank > 4) throw Error{ "HDF5: more than 4 dimensions" };
    if(p + 8 * (size_t) rank > sz) throw Error{ "HDF5: dataspace message too short" };
    v.rank = rank;
    for(int k = 0; k < rank; k++) v.dims[k] = get<u64>(d + p + 8 * k);
  }
  std::map<std::string, u64> members(u64 oh) const
  {
    std::map<std::string, u64> out;
    u64 bt = UNDEF, hp = UNDEF;
    for(const Message &q : header(oh))
      if(q.type == 0x0011 && q.size >= 16)
	{
	  bt = get<u64>(q.data);
	  hp = get<u64>(q.data + 8);
	}
    if(bt == UNDEF) throw Error{ "HDF5: object is not an (old-style) group" };
    const unsigned char *h = at(hp, 32);
    if(memcmp(h, "HEAP", 4)) throw Error{ "HDF5: local heap signature missing" };
    const u64 seg_size = get<u64>(h + 8), seg_addr = get<u64>(h + 24);
    const unsigned char *seg = at(seg_addr, seg_size);
    std::vector<u64> todo(1, bt);
    int guard = 0;
    while(!todo.empty())
      {
	if(++guard > 100000) throw Error{ "HDF5: group B-tree does not end" };
	const u64 a = todo.back();
	todo.pop_back();
	const unsigned char *t = at(a, 24);
	if(memcmp(t, "TREE", 4)) throw Error{ "HDF5: B-tree signature missing" };
	if(t[4] != 0) throw Error{ "HDF5: group B-tree expected" };
	const int level = t[5], used = get<unsigned short>(t + 6);
	const unsigned char *body = at(a + 24, (u64) (2 * used + 1) * 8);
	for(int k = 0; k < used; k++)
	  {
	    const u64 child = get<u64>(body + 8 + 16 * k);
	    if(level > 0)
	      todo.push_back(child);
	    else
	      {
		const unsigned char *s = at(child, 8);
		if(memcmp(s, "SNOD", 4)) throw Error{ "HDF5: symbol table node signature missing" };
		const int nsym = get<unsigned short>(s + 6);
		const unsigned char *e = at(child + 8, (u64) nsym * 40);
		for(int i = 0; i < nsym; i++)
		  {
		    const u64 no = get<u64>(e + 40 * i), obj = get<u64>(e + 40 * i + 8);
		    if(no >= seg_size) throw Error{ "HDF5: link name outside the heap" };
		    const void *z = memchr(seg + no, 0, (size_t) (seg_size - no));
		    if(!z) throw Error{ "HDF5: unterminated link name" };
		    out[std::string((const char *) seg + no)] = obj;
		  }
	      }
	  }
      }
    return out;
  }

public:
  Reader(const void *buf, u64 size) : b((const unsigned char *) buf), n(size)
  {
    const unsigned char sig[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };
    u64 off = 0;
    bool found = false;
    while(off + 96 <= n)
      {
	if(memcmp(b + off, sig, 8) == 0)
	  {
	    found = true;
	    break;
	  }
	off = off ? off * 2 : 512;
      }
    if(!found) throw Error{ "not an HDF5 file (no signature)" };
    const unsigned char *sb = b + off;
    const int ver = sb[8];
    if(ver != 0 && ver != 1) throw Error{ "HDF5: superblock version 2/3 (libver='latest') is not supported" };
    if(sb[13] != 8 || sb[14] != 8) throw Error{ "HDF5: only 8-byte offsets and lengths are supported" };
    const size_t p = 24 + (ver == 1 ? 4 : 0);
    base = get<u64>(sb + p);
    if(base != off && off != 0 && base == 0) base = off;	// a user block the base address does not mention
    root_oh = get<u64>(sb + p + 32 + 8);
  }
  // object header address of an absolute path ("/Header", "/PartType0/Coordinates"); UNDEF if a component is missing
  u64 find(const std::string &path) const
  {
    u64 cur = root_oh;
    size_t p = 0;
    while(p < path.size())
      {
	while(p < path.size() && path[p] == '/') p++;
	size_t q = path.find('/', p);
	if(q == std::string::npos) q = path.size();
	if(q == p) break;
	const std::map<std::string, u64> ms = members(cur);
	const auto it = ms.find(path.substr(p, q - p));
	if(it == ms.end()) return UNDEF;
	cur = it->second;
	p = q;
      }
    return cur;
  }
  std::vector<std::string> list(const std::string &path) const
  {
    const u64 oh = find(path);
    if(oh == UNDEF) throw Error{ "HDF5: no such group: " + path };
    std::vector<std::string> out;
    for(auto &kv : members(oh)) out.push_back(kv.first);
    return out;
  }
  bool attribute(u64 oh, const std::string &name, Value &v) const
  {
    for(const Message &q : header(oh))
      {
	if(q.type != 0x000C || q.size < 8) continue;
	const int ver = q.data[0];
	if(ver < 1 || ver > 3) throw Error{ "HDF5: unknown attribute version" };
	const size_t nsz = get<unsigned short>(q.data + 2), tsz = get<unsigned short>(q.data + 4), ssz = get<unsigned short>(q.data + 6);
	size_t p = 8 + (ver == 3 ? 1 : 0);
	auto pad = [&](size_t x) { return ver == 1 ? ((x + 7) & ~(size_t) 7) : x; };
	if(p + pad(nsz) + pad(tsz) + pad(ssz) > q.size) throw Error{ "HDF5: attribute message too short" };
	const std::string nm((const char *) q.data + p, strnlen((const char *) q.data + p, nsz));
	if(nm != name) continue;
	p += pad(nsz);
	parse_type(q.data + p, tsz, v);
	p += pad(tsz);
	parse_space(q.data + p, ssz, v);
	p += pad(ssz);
	if(p + v.count() * (u64) v.size > q.size) throw Error{ "HDF5: attribute data run past the message" };
	v.data = q.data + p;
	return true;
      }
    return false;
  }
  void dataset(u64 oh, Value &v) const
  {
    bool have_t = false, have_s = false, have_l = false;
    for(const Message &q : header(oh))
      {
	if(q.type == 0x0003)
	  {
	    parse_type(q.data, q.size, v);
	    have_t = true;
	  }
	else if(q.type == 0x0001)
	  {
	    parse_space(q.data, q.size, v);
	    have_s = true;
	  }
	else if(q.type == 0x000B)
	  throw Error{ "HDF5: filtered (compressed) datasets are not supported" };
      }
    if(!have_t || !have_s) throw Error{ "HDF5: object is not a dataset" };
    const u64 bytes = v.count() * (u64) v.size;
    for(const Message &q : header(oh))
      {
	if(q.type != 0x0008) continue;
	const unsigned char *d = q.data;
	const int ver = d[0];
	if(ver == 1 || ver == 2)
	  {
	    const int ndim = d[1], cls = d[2];
	    if(cls == 1)
	      v.data = bytes ? at(get<u64>(d + 8), bytes) : b;
	    else if(cls == 0)
	      {
		const size_t p = 8 + 4 * (size_t) ndim;
		if(p + 4 > q.size || p + 4 + get<unsigned int>(d + p) > q.size || get<unsigned int>(d + p) < bytes) throw Error{ "HDF5: compact data too short" };
		v.data = d + p + 4;
	      }
	    else
	      throw Error{ "HDF5: chunked datasets are not supported (write the file with contiguous layout)" };
	  }
	else if(ver == 3)
	  {
	    const int cls = d[1];
	    if(cls == 1)
	      {
		const u64 addr = get<u64>(d + 2), sz = get<u64>(d + 10);
		if(sz < bytes) throw Error{ "HDF5: dataset storage smaller than its dataspace" };
		if(bytes && addr == UNDEF) throw Error{ "HDF5: dataset has no storage (never written)" };
		v.data = bytes ? at(addr, bytes) : b;
	      }
	    else if(cls == 0)
	      {
		const size_t sz = get<unsigned short>(d + 2);
		if(4 + sz > q.size || sz < bytes) throw Error{ "HDF5: compact data too short" };
		v.data = d + 4;
	      }
	    else
	      throw Error{ "HDF5: chunked datasets are not supported (write the file with contiguous layout)" };
	  }
	else
	  throw Error{ "HDF5: unknown data layout version" };
	have_l = true;
      }
    if(!have_l) throw Error{ "HDF5: dataset without a layout message" };
  }
};
}	// namespace glh5
