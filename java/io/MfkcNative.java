package io;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

/**
 * Panama FFM (JDK >= 22) binding of libmfkc.so -- include/mfkc.h, one downcall handle per entry point.
 * NOT compiled in this repository (no JDK in the build image); it is the binding a MetaFast maintainer
 * adds next to src/io/IOUtils.java.  Every method is a 1:1 image of the C prototype.
 */
public final class MfkcNative {
    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("mfkc.library", "libmfkc.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), fd);
    }

    /** struct mfkc_cfg (88 bytes, see include/mfkc.h) */
    public static final StructLayout CFG = MemoryLayout.structLayout(
            JAVA_INT.withName("struct_size"), JAVA_INT.withName("k"), JAVA_INT.withName("min_seq_len"),
            JAVA_INT.withName("device"), JAVA_INT.withName("variant"), JAVA_INT.withName("n_shards"),
            JAVA_INT.withName("shard_id"), JAVA_INT.withName("reserved0"),
            JAVA_LONG.withName("table_slots"), JAVA_LONG.withName("expected_distinct"),
            JAVA_LONG.withName("max_table_bytes"), JAVA_LONG.withName("staging_bytes"),
            JAVA_INT.withName("region_shift"), JAVA_INT.withName("reserved2"),
            JAVA_LONG.withName("expected_kmers"), JAVA_LONG.withName("reserved1"));

    static final MethodHandle CREATE = h("mfkc_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle DESTROY = h("mfkc_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle LAST_ERROR = h("mfkc_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    static final MethodHandle RESET = h("mfkc_reset", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle PINNED_ALLOC = h("mfkc_pinned_alloc", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    static final MethodHandle SUBMIT_READS = h("mfkc_submit_reads", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle FLUSH = h("mfkc_flush", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle STATS = h("mfkc_stats", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle HISTOGRAM = h("mfkc_histogram", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle EMIT_BEGIN = h("mfkc_emit_begin", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle EMIT_NEXT = h("mfkc_emit_next", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    static final MethodHandle FC_LOAD = h("mfkc_fc_load_components", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle FC_SELECTED = h("mfkc_fc_set_selected", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle FC_RESET = h("mfkc_fc_reset_values", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle FC_ADD_RECORDS = h("mfkc_fc_add_records", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle FC_ADD_READS = h("mfkc_fc_add_reads", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle FC_FEATURES = h("mfkc_fc_features", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle READER_OPEN = h("mfkc_reader_open", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle READER_NEXT = h("mfkc_reader_next", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle READER_PENDING_BASES = h("mfkc_reader_pending_bases", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle READER_CLOSE = h("mfkc_reader_close", FunctionDescriptor.ofVoid(ADDRESS));

    // ---- the .kmers.bin consumers (kmers-filter, unique-kmers-multi, kmers-samples-counter, seq-builder): one mfkc_kset per
    // BigLong2ShortHashMap of src/tools/KmersFilter.java:94-110, UniqueKmersMultipleSamplesFinder.java:97-158,
    // KmersSamplesCounter.java:90-119, SeqBuilderMain.java:78-144
    static final MethodHandle KSET_CREATE = h("mfkc_kset_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle KSET_DESTROY = h("mfkc_kset_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle KSET_LOAD_RECORDS = h("mfkc_kset_load_records", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT));
    static final MethodHandle KSET_LOAD_FINISH = h("mfkc_kset_load_finish", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle KSET_SIZE = h("mfkc_kset_size", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle KSET_RESET_VALUES = h("mfkc_kset_reset_values", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle KSET_UPDATE = h("mfkc_kset_update", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
    static final MethodHandle KSET_SELECT_BEGIN = h("mfkc_kset_select_begin", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    static final MethodHandle KSET_SELECT_NEXT = h("mfkc_kset_select_next", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS));
    static final MethodHandle KSET_HISTOGRAM = h("mfkc_kset_histogram", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle KSET_SEQ_BEGIN = h("mfkc_kset_sequences_begin", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle KSET_SEQ_FETCH = h("mfkc_kset_sequences_fetch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    // ComponentsBuilder.splitStrategy(hm, k, b1, b2, ...) (src/algo/ComponentsBuilder.java:24-31) on a device map
    static final MethodHandle KSET_COMP_BEGIN = h("mfkc_kset_components_begin", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
    static final MethodHandle KSET_COMP_FETCH = h("mfkc_kset_components_fetch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));

    // ---- multi-GPU: one context per GPU (cfg.n_shards / shard_id); a single JVM attaches the contexts to each other directly
    static final MethodHandle P2P_STAGE_CREATE = h("mfkc_p2p_stage_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG));
    // the bin-local flavour (k <= 31): geometry from the expected k-mers per GPU, (owner, bin) staging + overflow pool
    static final MethodHandle P2P_BIN_GEOMETRY = h("mfkc_p2p_bin_geometry", FunctionDescriptor.of(JAVA_INT, JAVA_LONG, JAVA_INT, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle P2P_STAGE_CREATE_BINS = h("mfkc_p2p_stage_create_bins", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG));
    static final MethodHandle MERGE_RECORDS = h("mfkc_merge_records", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle FC_ADD_EMITTED = h("mfkc_fc_add_emitted", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle LIBRARY_NAME = h("mfkc_library_name", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    static final MethodHandle P2P_ATTACH_CTX = h("mfkc_p2p_attach_ctx", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle P2P_STAGE_RESET = h("mfkc_p2p_stage_reset", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle P2P_SUBMIT_READS = h("mfkc_p2p_submit_reads", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle P2P_COUNTS = h("mfkc_p2p_counts", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle P2P_DRAIN = h("mfkc_p2p_drain", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));

    /** rc != 0 -> ExecutionFailedException with mfkc_last_error(ctx), like the tools' IOException wrapping (Tool.java:286-287). */
    static void check(MemorySegment ctx, int rc) throws ru.ifmo.genetics.utils.tool.ExecutionFailedException {
        if (rc == 0) return;
        try {
            MemorySegment msg = (MemorySegment) LAST_ERROR.invokeExact(ctx);
            throw new ru.ifmo.genetics.utils.tool.ExecutionFailedException(
                    "libmfkc error " + rc + ": " + msg.reinterpret(4096).getString(0));
        } catch (ru.ifmo.genetics.utils.tool.ExecutionFailedException e) {
            throw e;
        } catch (Throwable t) {
            throw new ru.ifmo.genetics.utils.tool.ExecutionFailedException("libmfkc error " + rc, t);
        }
    }

    private MfkcNative() {}
}
