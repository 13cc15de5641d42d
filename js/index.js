// js/index.js — the reference's plugin surface (src/index.js:16-31) over the N-API shim (js/napi/shim.cc).
//
// NOT RUNNABLE IN THIS IMAGE (no Node, no GL; SURVEY.md D5) — this is the reference-side glue a maintainer would
// add to drop the CUDA engine under a postprocessing.js EffectComposer: same class names, constructor signatures,
// option names/defaults and update/setSize/reset/dispose methods as the reference.  The tested host mirror of the
// same surface is realism_effects_b200/effects.py.
//
// Plane transport in this first binding is host memory: the G-buffer / velocity / scene-colour render targets are
// read back (renderer.readRenderTargetPixels) into typed arrays and passed to rfx_ssgi_chain_render_host, which
// uploads, renders and downloads `composed`.  The zero-copy path (cudaGraphicsGLRegisterImage on the WebGL
// textures) is SURVEY.md §8(f) row 4.
import { createRequire } from "node:module"
const rfx = createRequire(import.meta.url)("./napi/rfx_napi.node")

// src/ssgi/SSGIOptions.js:26-48
export const defaultSSGIOptions = {
	mode: "ssgi", distance: 10, thickness: 10, denoiseIterations: 1, denoiseKernel: 2, denoiseDiffuse: 10, denoiseSpecular: 10,
	radius: 3, phi: 0.5, lumaPhi: 5, depthPhi: 2, normalPhi: 50, roughnessPhi: 50, specularPhi: 50, envBlur: 0.5,
	importanceSampling: true, steps: 20, refineSteps: 5, resolutionScale: 1, missedRays: false, outputTexture: null
}

const FLAG = { importanceSampling: 1, missedRays: 2, useDirectLight: 4, useEnvMap: 8 }
let sharedCtx = null
const context = (device = 0) => (sharedCtx ??= rfx.ctxCreate(device))

const f32 = m => new Float32Array(m.elements ?? m)
const cameraBlock = camera => ({
	projection: f32(camera.projectionMatrix), projectionInverse: f32(camera.projectionMatrixInverse),
	matrixWorld: f32(camera.matrixWorld), matrixWorldInverse: f32(camera.matrixWorldInverse), near: camera.near, far: camera.far
})

export class VelocityDepthNormalPass {
	// new VelocityDepthNormalPass(scene, camera) — src/temporal-reproject/pass/VelocityDepthNormalPass.js:71-91.
	// Rasterisation stays in three.js; this wrapper only exposes the planes the CUDA engine consumes.
	constructor(scene, camera, rasterPass) { this._scene = scene; this._camera = camera; this.rasterPass = rasterPass; this.needsSwap = false }
	get texture() { return this.rasterPass.texture }
	get renderTarget() { return this.rasterPass.renderTarget }
	setSize(w, h) { this.rasterPass.setSize(w, h) }
	render(renderer) { this.rasterPass.render(renderer) }
	dispose() { this.rasterPass.dispose() }
}

export class SSGIEffect {
	// new SSGIEffect(composer, scene, camera, options) — src/ssgi/SSGIEffect.js:31 (the code's signature, SURVEY.md D6)
	constructor(composer, scene, camera, options = {}) {
		this.composer = composer; this._scene = scene; this._camera = camera
		const opts = { ...defaultSSGIOptions, ...options }
		this._options = opts
		this.ctx = context(options.device ?? 0)
		this.velocityDepthNormalPass = options.velocityDepthNormalPass
		this.lastCamera = null
		for (const key of Object.keys(opts)) {
			Object.defineProperty(this, key, {
				get: () => opts[key],
				set: value => { if (opts[key] === value) return; opts[key] = value; this._rebuild() }   // setters end with reset()
			})
		}
		this.setSize(options.width, options.height)
	}
	_flags() {
		const o = this._options
		return (o.importanceSampling && this._hasEnv ? FLAG.importanceSampling : 0) | (o.missedRays ? FLAG.missedRays : 0) |
			FLAG.useDirectLight | (this._hasEnv ? FLAG.useEnvMap : 0)
	}
	_rebuild() {
		if (this.chain) rfx.chainDestroy(this.chain)
		this.chain = rfx.chainCreate(this.ctx, { ...this._options, width: this.width, height: this.height, flags: this._flags(), mode: this._options.mode === "ssr" ? 1 : 0 })
		this.out = new Float32Array(this.width * this.height * 4)
	}
	setSize(width, height) { if (width === undefined || (width === this.width && height === this.height)) return; this.width = width; this.height = height; this._rebuild() }
	reset() { rfx.chainReset(this.chain) }
	get outputTexture() { return this.out }
	// update(renderer, inputBuffer) — src/ssgi/SSGIEffect.js:372-404.  `planes` = typed arrays read back from the
	// GBufferPass / VelocityDepthNormalPass / input-buffer render targets by the caller's glue.
	update(renderer, inputBuffer, deltaTime, planes) {
		const cam = cameraBlock(this._camera)
		const moved = !this.lastCamera || cam.matrixWorld.some((v, i) => Math.abs(v - this.lastCamera[i]) > 1e-6)
		this.lastCamera = cam.matrixWorld
		rfx.chainRenderHost(this.ctx, this.chain, cam, planes.depth, planes.gbuffer, planes.velocity, planes.directLight ?? null,
			new Float32Array(this._camera.position.toArray()), moved, this.out)
	}
	dispose() { if (this.chain) rfx.chainDestroy(this.chain); this.chain = null }
}
SSGIEffect.DefaultOptions = defaultSSGIOptions

export class SSREffect extends SSGIEffect {
	// src/ssgi/SSREffect.js:3-9
	constructor(composer, scene, camera, options = {}) { super(composer, scene, camera, { ...options, mode: "ssr" }) }
}

// TRAAEffect / MotionBlurEffect / HBAOEffect / TemporalReprojectPass / PoissonDenoisePass bind the per-pass entry points
// (rfx_temporal_reproject_launch, rfx_motion_blur_launch, rfx_hbao_launch, rfx_ao_compose_launch,
// rfx_poisson_denoise_launch) the same way; their tested host mirror is realism_effects_b200/effects.py.
