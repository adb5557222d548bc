/* Prototype-only shim for the system libsqlite3.so.0 (no sqlite3.h in this image).
 * TEST INFRASTRUCTURE: used only to compile the UNMODIFIED reference sources from
 * /root/reference/src into oracle/_ref/ (see oracle/build_ref.sh).  Declares exactly
 * the subset of the public sqlite3 C API the reference calls. */
#ifndef FXO_SQLITE3_SHIM_H
#define FXO_SQLITE3_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
#define SQLITE_VERSION "3.45.1-system"
#define SQLITE_OK     0
#define SQLITE_ERROR  1
#define SQLITE_ROW    100
#define SQLITE_DONE   101
typedef struct sqlite3 sqlite3;
typedef struct sqlite3_stmt sqlite3_stmt;
typedef long long sqlite3_int64;
typedef void (*sqlite3_destructor_type)(void *);
#define SQLITE_STATIC    ((sqlite3_destructor_type)0)
#define SQLITE_TRANSIENT ((sqlite3_destructor_type)-1)
int sqlite3_open(const char *filename, sqlite3 **db);
int sqlite3_close(sqlite3 *db);
int sqlite3_exec(sqlite3 *db, const char *sql, int (*cb)(void *, int, char **, char **), void *arg, char **errmsg);
int sqlite3_prepare_v2(sqlite3 *db, const char *sql, int nbyte, sqlite3_stmt **stmt, const char **tail);
int sqlite3_step(sqlite3_stmt *stmt);
int sqlite3_reset(sqlite3_stmt *stmt);
int sqlite3_finalize(sqlite3_stmt *stmt);
int sqlite3_bind_null(sqlite3_stmt *stmt, int idx);
int sqlite3_bind_int(sqlite3_stmt *stmt, int idx, int v);
int sqlite3_bind_int64(sqlite3_stmt *stmt, int idx, sqlite3_int64 v);
int sqlite3_bind_double(sqlite3_stmt *stmt, int idx, double v);
int sqlite3_bind_text(sqlite3_stmt *stmt, int idx, const char *s, int n, sqlite3_destructor_type d);
int sqlite3_bind_blob(sqlite3_stmt *stmt, int idx, const void *p, int n, sqlite3_destructor_type d);
int sqlite3_column_int(sqlite3_stmt *stmt, int col);
sqlite3_int64 sqlite3_column_int64(sqlite3_stmt *stmt, int col);
double sqlite3_column_double(sqlite3_stmt *stmt, int col);
const unsigned char *sqlite3_column_text(sqlite3_stmt *stmt, int col);
const void *sqlite3_column_blob(sqlite3_stmt *stmt, int col);
int sqlite3_column_bytes(sqlite3_stmt *stmt, int col);
char *sqlite3_mprintf(const char *fmt, ...);
void sqlite3_free(void *p);
const char *sqlite3_errmsg(sqlite3 *db);
#ifdef __cplusplus
}
#endif
#endif
