// SQLite-backed Database with the reference's schema and blob formats, written against the public SQLite C API
// resolved at run time (dlopen), see Database.h.
#include "Database/Database.h"

#include <dlfcn.h>

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace MonocularSfM;

namespace {
// ---- the small part of the SQLite C API that is used (public, stable ABI)
struct sqlite3;
struct sqlite3_stmt;
typedef long long sqlite3_int64;
const int kOk = 0, kRow = 100, kDone = 101;
const int kOpenReadWrite = 0x2, kOpenCreate = 0x4, kOpenNoMutex = 0x8000;
typedef void (*destructor_t)(void*);
#define SQLITE_TRANSIENT_ ((destructor_t)-1)

struct Api {
    int (*open_v2)(const char*, sqlite3**, int, const char*);
    int (*close)(sqlite3*);
    int (*exec)(sqlite3*, const char*, int (*)(void*, int, char**, char**), void*, char**);
    int (*prepare_v2)(sqlite3*, const char*, int, sqlite3_stmt**, const char**);
    int (*finalize)(sqlite3_stmt*);
    int (*step)(sqlite3_stmt*);
    int (*reset)(sqlite3_stmt*);
    int (*bind_int64)(sqlite3_stmt*, int, sqlite3_int64);
    int (*bind_blob)(sqlite3_stmt*, int, const void*, int, destructor_t);
    int (*bind_text)(sqlite3_stmt*, int, const char*, int, destructor_t);
    sqlite3_int64 (*column_int64)(sqlite3_stmt*, int);
    const void* (*column_blob)(sqlite3_stmt*, int);
    int (*column_bytes)(sqlite3_stmt*, int);
    const unsigned char* (*column_text)(sqlite3_stmt*, int);
    sqlite3_int64 (*last_insert_rowid)(sqlite3*);
    const char* (*errmsg)(sqlite3*);
    void (*free_)(void*);
};

const Api& api() {
    static Api a;
    static bool loaded = false;
    if (loaded) return a;
    void* lib = dlopen("libsqlite3.so.0", RTLD_NOW);
    if (!lib) lib = dlopen("libsqlite3.so", RTLD_NOW);
    if (!lib) {
        std::fprintf(stderr, "SQLite error: cannot load libsqlite3.so.0: %s\n", dlerror());
        std::exit(EXIT_FAILURE);
    }
#define LOAD(field, name)                                                        \
    *reinterpret_cast<void**>(&a.field) = dlsym(lib, name);                      \
    if (!a.field) { std::fprintf(stderr, "SQLite error: missing symbol %s\n", name); std::exit(EXIT_FAILURE); }
    LOAD(open_v2, "sqlite3_open_v2") LOAD(close, "sqlite3_close") LOAD(exec, "sqlite3_exec")
    LOAD(prepare_v2, "sqlite3_prepare_v2") LOAD(finalize, "sqlite3_finalize") LOAD(step, "sqlite3_step")
    LOAD(reset, "sqlite3_reset") LOAD(bind_int64, "sqlite3_bind_int64") LOAD(bind_blob, "sqlite3_bind_blob")
    LOAD(bind_text, "sqlite3_bind_text") LOAD(column_int64, "sqlite3_column_int64") LOAD(column_blob, "sqlite3_column_blob")
    LOAD(column_bytes, "sqlite3_column_bytes") LOAD(column_text, "sqlite3_column_text")
    LOAD(last_insert_rowid, "sqlite3_last_insert_rowid") LOAD(errmsg, "sqlite3_errmsg") LOAD(free_, "sqlite3_free")
#undef LOAD
    loaded = true;
    return a;
}

const size_t kMaxNumImages = 10000;   // Database.cpp:6 — pair ids are 10000 * min + max
}  // namespace

struct Database::Impl {
    sqlite3* db = nullptr;
    // number of rows of the blob stored under `key` (0 when absent)
    size_t blob_rows(const char* sql, sqlite3_int64 key) const {
        sqlite3_stmt* st = prepare(sql);
        check(api().bind_int64(st, 1, key), __LINE__);
        size_t n = 0;
        if (check(api().step(st), __LINE__) == kRow) n = static_cast<size_t>(api().column_int64(st, 0));
        api().finalize(st);
        return n;
    }
    Database::Image image_row(const char* sql, const image_t* id, const std::string* name) const {
        sqlite3_stmt* st = prepare(sql);
        if (id) check(api().bind_int64(st, 1, *id), __LINE__);
        else check(api().bind_text(st, 1, name->c_str(), static_cast<int>(name->size()), SQLITE_TRANSIENT_), __LINE__);
        Database::Image im;
        im.id = INVALID;
        if (check(api().step(st), __LINE__) == kRow) {
            im.id = static_cast<image_t>(api().column_int64(st, 0));
            im.name = reinterpret_cast<const char*>(api().column_text(st, 1));
        }
        api().finalize(st);
        return im;
    }
    // Failures are fatal like in the reference (Database.cpp:8-22): message on stderr, exit(EXIT_FAILURE).
    int check(int rc, int line) const {
        if (rc == kOk || rc == kRow || rc == kDone) return rc;
        std::fprintf(stderr, "SQLite error [Database.cpp, line %d]: %s\n", line, db ? api().errmsg(db) : "?");
        std::exit(EXIT_FAILURE);
    }
    void exec(const char* sql) const {
        char* err = nullptr;
        if (api().exec(db, sql, nullptr, nullptr, &err) != kOk) {
            std::fprintf(stderr, "SQLite error: %s (%s)\n", err ? err : "?", sql);   // the reference only prints (:26-36)
            if (err) api().free_(err);
        }
    }
    sqlite3_stmt* prepare(const char* sql) const {
        sqlite3_stmt* st = nullptr;
        check(api().prepare_v2(db, sql, -1, &st, nullptr), __LINE__);
        return st;
    }
    bool exists(const char* sql, sqlite3_int64 key) const {
        sqlite3_stmt* st = prepare(sql);
        check(api().bind_int64(st, 1, key), __LINE__);
        const bool found = check(api().step(st), __LINE__) == kRow;
        api().finalize(st);
        return found;
    }
    // rows x cols blob of T at columns (c, c+1, c+2) = rows, cols, data
    template <class T>
    bool read_blob(const char* sql, sqlite3_int64 key, std::vector<T>& out, size_t& rows, size_t& cols) const {
        sqlite3_stmt* st = prepare(sql);
        check(api().bind_int64(st, 1, key), __LINE__);
        rows = cols = 0;
        out.clear();
        const bool found = check(api().step(st), __LINE__) == kRow;
        if (found) {
            rows = static_cast<size_t>(api().column_int64(st, 0));
            cols = static_cast<size_t>(api().column_int64(st, 1));
            const size_t bytes = static_cast<size_t>(api().column_bytes(st, 2));
            assert(bytes == rows * cols * sizeof(T));
            out.resize(rows * cols);
            if (bytes) std::memcpy(out.data(), api().column_blob(st, 2), bytes);
        }
        api().finalize(st);
        return found;
    }
    template <class T>
    void write_blob(const char* sql, sqlite3_int64 key, const std::vector<T>& data, size_t rows, size_t cols) const {
        sqlite3_stmt* st = prepare(sql);
        check(api().bind_int64(st, 1, key), __LINE__);
        check(api().bind_int64(st, 2, static_cast<sqlite3_int64>(rows)), __LINE__);
        check(api().bind_int64(st, 3, static_cast<sqlite3_int64>(cols)), __LINE__);
        check(api().bind_blob(st, 4, data.empty() ? static_cast<const void*>("") : data.data(),
                              static_cast<int>(data.size() * sizeof(T)), SQLITE_TRANSIENT_), __LINE__);
        check(api().step(st), __LINE__);
        api().finalize(st);
    }
};

Database::Database() : impl_(new Impl()) {}
Database::~Database() {
    Close();
    delete impl_;
}

void Database::Open(const std::string& path) {
    impl_->check(api().open_v2(path.c_str(), &impl_->db, kOpenCreate | kOpenReadWrite | kOpenNoMutex, nullptr), __LINE__);
    // same pragmas as Database.cpp:299-302
    impl_->exec("PRAGMA synchronous=OFF");
    impl_->exec("PRAGMA journal_mode=WAL");
    impl_->exec("PRAGMA temp_store=MEMORY");
    impl_->exec("PRAGMA foreign_keys=ON");
    // schema of Database.cpp:710-764
    impl_->exec("CREATE TABLE IF NOT EXISTS images(image_id INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, name TEXT NOT NULL UNIQUE)");
    const char* blob_tables[] = {"keypoints", "colors", "descriptors"};
    for (const char* t : blob_tables) {
        const std::string sql = std::string("CREATE TABLE IF NOT EXISTS ") + t +
                                "(image_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB, "
                                "FOREIGN KEY(image_id) REFERENCES images(image_id) ON DELETE CASCADE)";
        impl_->exec(sql.c_str());
    }
    impl_->exec("CREATE TABLE IF NOT EXISTS matches(pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB)");
    impl_->exec(("PRAGMA user_version=" + std::to_string(kSchemaVersion)).c_str());
}

void Database::Close() {
    if (impl_->db) {
        api().close(impl_->db);
        impl_->db = nullptr;
    }
}

void Database::BeginTransaction() const { impl_->exec("BEGIN TRANSACTION"); }
void Database::EndTransaction() const { impl_->exec("END TRANSACTION"); }

bool Database::ExistImageById(const image_t id) const { return impl_->exists("SELECT 1 FROM images WHERE image_id = ?", id); }
bool Database::ExistImageByName(const std::string name) const { return ReadImageByName(name).id != INVALID; }
bool Database::ExistKeyPointsColor(const image_t id) const { return impl_->exists("SELECT 1 FROM colors WHERE image_id = ?", id); }
bool Database::ExistKeyPoints(const image_t id) const { return impl_->exists("SELECT 1 FROM keypoints WHERE image_id = ?", id); }
bool Database::ExistDescriptors(const image_t id) const { return impl_->exists("SELECT 1 FROM descriptors WHERE image_id = ?", id); }
bool Database::ExistMatches(const image_pair_t pair_id) const { return impl_->exists("SELECT 1 FROM matches WHERE pair_id = ?", pair_id); }
bool Database::ExistMatches(const image_t a, const image_t b) const { return ExistMatches(ImagePairToPairId(a, b)); }

size_t Database::NumImages() const {
    sqlite3_stmt* st = impl_->prepare("SELECT COUNT(*) FROM images");
    impl_->check(api().step(st), __LINE__);
    const size_t n = static_cast<size_t>(api().column_int64(st, 0));
    api().finalize(st);
    return n;
}
size_t Database::NumKeyPoints(const image_t id) const { return impl_->blob_rows("SELECT rows FROM keypoints WHERE image_id = ?", id); }
size_t Database::NumKeyPointsColor(const image_t id) const { return impl_->blob_rows("SELECT rows FROM colors WHERE image_id = ?", id); }
size_t Database::NumMatches(const image_pair_t pair_id) const { return impl_->blob_rows("SELECT rows FROM matches WHERE pair_id = ?", pair_id); }
Database::Image Database::ReadImageById(const image_t id) const { return impl_->image_row("SELECT image_id, name FROM images WHERE image_id = ?", &id, nullptr); }
Database::Image Database::ReadImageByName(const std::string name) const { return impl_->image_row("SELECT image_id, name FROM images WHERE name = ?", nullptr, &name); }
size_t Database::NumDescriptors(const image_t id) const {
    sqlite3_stmt* st = impl_->prepare("SELECT rows FROM descriptors WHERE image_id = ?");
    impl_->check(api().bind_int64(st, 1, id), __LINE__);
    size_t n = 0;
    if (impl_->check(api().step(st), __LINE__) == kRow) n = static_cast<size_t>(api().column_int64(st, 0));
    api().finalize(st);
    return n;
}
size_t Database::NumMatches(const image_t a, const image_t b) const {
    sqlite3_stmt* st = impl_->prepare("SELECT rows FROM matches WHERE pair_id = ?");
    impl_->check(api().bind_int64(st, 1, ImagePairToPairId(a, b)), __LINE__);
    size_t n = 0;
    if (impl_->check(api().step(st), __LINE__) == kRow) n = static_cast<size_t>(api().column_int64(st, 0));
    api().finalize(st);
    return n;
}

std::vector<Database::Image> Database::ReadAllImages() const {
    std::vector<Image> out;
    sqlite3_stmt* st = impl_->prepare("SELECT image_id, name FROM images ORDER BY image_id");
    while (impl_->check(api().step(st), __LINE__) == kRow) {
        Image im;
        im.id = static_cast<image_t>(api().column_int64(st, 0));
        im.name = reinterpret_cast<const char*>(api().column_text(st, 1));
        out.push_back(im);
    }
    api().finalize(st);
    return out;
}

image_t Database::WriteImage(const Image& image, const bool use_image_id) const {
    sqlite3_stmt* st = impl_->prepare("INSERT INTO images(image_id, name) VALUES(?, ?)");
    if (use_image_id) impl_->check(api().bind_int64(st, 1, image.id), __LINE__);   // else NULL -> AUTOINCREMENT
    impl_->check(api().bind_text(st, 2, image.name.c_str(), static_cast<int>(image.name.size()), SQLITE_TRANSIENT_), __LINE__);
    impl_->check(api().step(st), __LINE__);
    api().finalize(st);
    return static_cast<image_t>(api().last_insert_rowid(impl_->db));
}

// keypoints blob: rows x 4 float32 (x, y, size, angle) — KeyPointsToBlob, Database.cpp
void Database::WriteKeyPoints(const image_t id, const std::vector<cv::KeyPoint>& kps) const {
    std::vector<float> v(kps.size() * 4);
    for (size_t i = 0; i < kps.size(); ++i) {
        v[4 * i] = kps[i].pt.x; v[4 * i + 1] = kps[i].pt.y; v[4 * i + 2] = kps[i].size; v[4 * i + 3] = kps[i].angle;
    }
    impl_->write_blob("INSERT INTO keypoints(image_id, rows, cols, data) VALUES(?, ?, ?, ?)", id, v, kps.size(), 4);
}
std::vector<cv::KeyPoint> Database::ReadKeyPoints(const image_t id) const {
    std::vector<float> v;
    size_t rows, cols;
    impl_->read_blob("SELECT rows, cols, data FROM keypoints WHERE image_id = ?", id, v, rows, cols);
    std::vector<cv::KeyPoint> out(rows);
    for (size_t i = 0; i < rows; ++i) {
        out[i].pt.x = v[i * cols];
        out[i].pt.y = v[i * cols + 1];
        if (cols > 2) out[i].size = v[i * cols + 2];
        if (cols > 3) out[i].angle = v[i * cols + 3];
    }
    return out;
}

// colors blob: rows x 3 uint8 (KeyPointsColorToBlob, Database.cpp:143-155)
void Database::WriteKeyPointsColor(const image_t id, const std::vector<cv::Vec3b>& colors) const {
    std::vector<unsigned char> v(colors.size() * 3);
    for (size_t i = 0; i < colors.size(); ++i)
        for (int k = 0; k < 3; ++k) v[3 * i + k] = colors[i][k];
    impl_->write_blob("INSERT INTO colors(image_id, rows, cols, data) VALUES(?, ?, ?, ?)", id, v, colors.size(), 3);
}
std::vector<cv::Vec3b> Database::ReadKeyPointsColor(const image_t id) const {
    std::vector<unsigned char> v;
    size_t rows, cols;
    impl_->read_blob("SELECT rows, cols, data FROM colors WHERE image_id = ?", id, v, rows, cols);
    assert(rows == 0 || cols == 3);
    std::vector<cv::Vec3b> out(rows);
    for (size_t i = 0; i < rows; ++i)
        for (int k = 0; k < 3; ++k) out[i][k] = v[3 * i + k];
    return out;
}

void Database::WriteDescriptors(const image_t id, const cv::Mat& desc) const {
    assert(desc.type() == CV_32F);   // Database.cpp:176
    std::vector<float> v(static_cast<size_t>(desc.rows) * desc.cols);
    for (int i = 0; i < desc.rows; ++i) std::memcpy(v.data() + static_cast<size_t>(i) * desc.cols, desc.ptr<float>(i), desc.cols * sizeof(float));
    impl_->write_blob("INSERT INTO descriptors(image_id, rows, cols, data) VALUES(?, ?, ?, ?)", id, v, desc.rows, desc.cols);
}
cv::Mat Database::ReadDescriptors(const image_t id) const {
    std::vector<float> v;
    size_t rows, cols;
    impl_->read_blob("SELECT rows, cols, data FROM descriptors WHERE image_id = ?", id, v, rows, cols);
    cv::Mat m(static_cast<int>(rows), static_cast<int>(cols), CV_32F);
    for (size_t i = 0; i < rows; ++i) std::memcpy(m.ptr<float>(static_cast<int>(i)), v.data() + i * cols, cols * sizeof(float));
    return m;
}

void Database::WriteMatches(const image_t a, const image_t b, const std::vector<cv::DMatch>& matches) const {
    const bool swap = SwapImagePair(a, b);     // stored with image_id1 < image_id2 orientation (Database.cpp:637-640)
    std::vector<int32_t> v(matches.size() * 2);
    for (size_t i = 0; i < matches.size(); ++i) {
        v[2 * i] = swap ? matches[i].trainIdx : matches[i].queryIdx;
        v[2 * i + 1] = swap ? matches[i].queryIdx : matches[i].trainIdx;
    }
    impl_->write_blob("INSERT INTO matches(pair_id, rows, cols, data) VALUES(?, ?, ?, ?)", ImagePairToPairId(a, b), v, matches.size(), 2);
}
std::vector<cv::DMatch> Database::ReadMatches(const image_t a, const image_t b) const {
    std::vector<int32_t> v;
    size_t rows, cols;
    impl_->read_blob("SELECT rows, cols, data FROM matches WHERE pair_id = ?", ImagePairToPairId(a, b), v, rows, cols);
    const bool swap = SwapImagePair(a, b);
    std::vector<cv::DMatch> out(rows);
    for (size_t i = 0; i < rows; ++i) {
        out[i].queryIdx = swap ? v[2 * i + 1] : v[2 * i];
        out[i].trainIdx = swap ? v[2 * i] : v[2 * i + 1];
    }
    return out;
}
std::vector<cv::DMatch> Database::ReadMatches(const image_pair_t pair_id) const {
    image_t a, b;
    PairIdToImagePair(pair_id, &a, &b);          // a < b: the stored orientation (Database.cpp:524-534)
    return ReadMatches(a, b);
}
std::vector<std::pair<image_pair_t, std::vector<cv::DMatch>>> Database::ReadAllMatches() const {
    std::vector<std::pair<image_pair_t, std::vector<cv::DMatch>>> out;
    sqlite3_stmt* st = impl_->prepare("SELECT pair_id, rows, cols, data FROM matches WHERE rows > 0");   // as the reference's statement
    while (impl_->check(api().step(st), __LINE__) == kRow) {
        const image_pair_t pid = static_cast<image_pair_t>(api().column_int64(st, 0));
        const size_t rows = static_cast<size_t>(api().column_int64(st, 1));
        const int32_t* d = static_cast<const int32_t*>(api().column_blob(st, 3));
        std::vector<cv::DMatch> m(rows);
        for (size_t i = 0; i < rows; ++i) { m[i].queryIdx = d[2 * i]; m[i].trainIdx = d[2 * i + 1]; }
        out.emplace_back(pid, std::move(m));
    }
    api().finalize(st);
    return out;
}

image_pair_t Database::ImagePairToPairId(const image_t a, const image_t b) {
    assert(a >= 0 && b >= 0 && static_cast<size_t>(a) < kMaxNumImages && static_cast<size_t>(b) < kMaxNumImages);
    return SwapImagePair(a, b) ? static_cast<image_pair_t>(kMaxNumImages * b + a) : static_cast<image_pair_t>(kMaxNumImages * a + b);
}
void Database::PairIdToImagePair(const image_pair_t pair_id, image_t* a, image_t* b) {
    *b = static_cast<image_t>(pair_id % kMaxNumImages);
    *a = static_cast<image_t>((pair_id - *b) / kMaxNumImages);
}
bool Database::SwapImagePair(const image_t a, const image_t b) { return a > b; }
